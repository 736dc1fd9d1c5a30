// coarse_grid.cu -- the whole coarse-grid PCG solve in ONE kernel on ALL SMs.
//
// coarse_cluster.cu keeps the solve on one 16-CTA cluster: right for one rank's 9 261 unknowns, but the replicated
// coarse problem of an 8-GPU job (68 921 unknowns, a 15 MB ELL matrix) is then read by 16 SMs through L2 every
// iteration: 850 us per coarse solve against 158 us on one GPU, the largest single loss of BPS5 weak scaling
// (DESIGN.md section 5).  Here the rows are spread over up to 148 co-resident CTAs (cooperative launch):
//   * each CTA keeps its ELL slice in shared memory (<= 512 rows x 27 entries x 8 B) and the six Krylov vectors of its
//     rows in registers;
//   * the SpMV input goes through a double-buffered L2-resident vector; two grid barriers per iteration (fence +
//     arrival counter + poll; the new u must be visible before the SpMV, the inner products before the update)
//     replace the kernel boundaries.  Measured (68 921 rows, 16 iterations, one B200): 151 us, against 481 us for the
//     cluster kernel and 295 us for the multi-launch path.  Flag-in-data words {value | epoch} instead of the
//     barriers (the protocol of the halo exchange, gs.cu) were slower here: 172 us with flagged inner-product
//     partials only, 309 us with the SpMV input flagged as well -- 35 000 threads polling their 27 gather targets
//     slow down the very stores they wait for;
//   * the partials of all CTAs are folded in a fixed order by every CTA (same bits everywhere), so alpha, beta and
//     the convergence test need no broadcast.
// Same recurrences, SpMV summation order (ascending ELL slot) and `checkEvery` test as coarse.cu / coarse_cluster.cu;
// the only difference is the grouping of the inner-product partial sums (per CTA of this grid).
#include "host.hpp"

namespace nrsb {

namespace {

constexpr int kGThreads = 512;

struct GridArgs {
  int NT, W, RPC, NTpad;
  const int* cols;
  const float* vals;
  const float* invDiag;
  const float* weight;
  const int* rowNode;
  const int* tIndex;
  const float* rhsE;
  float* xE;
  long Nlocal;
  int maxIter, checkEvery;
  double tol2;
  int matInSmem;
  double* S;         // [0]=gamma, [1]=gamma0, [2]=iterations
  float* uG;                 // [2][NTpad]
  double* redG;              // [2][gridDim][2]
  unsigned* bar;             // zero at launch
  int* err;
  // replicated multi-rank solve: right-hand-side all-gather through peer windows (as coarse_cluster.cu)
  int nranks, myRank;
  int nOwn;
  const int* ownG;
  const int* ownNode;
  float* const* peerWin;
  unsigned long long* const* peerFlags;
  const float* myWin;
  const unsigned long long* myFlags;
  unsigned long long epoch;
};

__device__ __forceinline__ double warp_sum_dd(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (co-resident) grid; `target` = number of arrivals after this barrier.  Arrival is one
// red.release (the CTA's stores, ordered before it by the block barrier, become visible first: cumulativity; no
// return value to wait for), the poll is relaxed with one acquire load at the end (every poll of an acquire load
// also invalidates L1: LDG.STRONG + CCTL.IVALL).
__device__ __forceinline__ void grid_barrier(unsigned* bar, const unsigned target, int* err)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const long long t0 = clock64();
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (clock64() - t0 > (1ll << 33)) {  // ~4 s: the grid is not co-resident after all
        if (err) *err = 1;
        break;
      }
    } while (v < target);
    (void)ld_acquire_u32(bar);
  }
  __syncthreads();
}

template <int RMAX>
__global__ void __launch_bounds__(kGThreads, 1) coarse_pcg_grid_kernel(const GridArgs a)
{
  const int C = (int)gridDim.x;
  const int c = (int)blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned nbar = 0;

  extern __shared__ __align__(16) unsigned char smraw[];
  double* wred = reinterpret_cast<double*>(smraw);  // [16 warps][2]
  double* tot = wred + 2 * (kGThreads / 32);        // [2]
  float* mvals = reinterpret_cast<float*>(tot + 2);  // [W][RPC]
  int* mcols = reinterpret_cast<int*>(mvals + (size_t)a.W * a.RPC);

  const int row0 = c * a.RPC;
  const int nRows = max(0, min(a.RPC, a.NT - row0));

  const float* mv;
  const int* mc;
  int mstride, moff;
  if (a.matInSmem) {
    for (int k = 0; k < a.W; ++k)
      for (int lr = tid; lr < nRows; lr += kGThreads) {
        mvals[k * a.RPC + lr] = a.vals[(size_t)k * a.NT + row0 + lr];
        mcols[k * a.RPC + lr] = a.cols[(size_t)k * a.NT + row0 + lr];
      }
    mv = mvals;
    mc = mcols;
    mstride = a.RPC;
    moff = 0;
  } else {
    mv = a.vals;
    mc = a.cols;
    mstride = a.NT;
    moff = row0;
  }

  // several ranks: all-gather the right-hand side (every rank pushes what it owns to everybody)
  const float* bsrc = nullptr;
  if (a.nranks > 1) {
    const size_t woff = (size_t)(a.epoch & 1ull) * a.NTpad;
    for (int i = c * kGThreads + tid; i < a.nOwn; i += C * kGThreads) {
      const float v = a.rhsE[a.ownNode[i]];
      const int g = a.ownG[i];
      for (int pr = 0; pr < a.nranks; ++pr) a.peerWin[pr][woff + g] = v;  // NVLink stores (own window included)
    }
    __threadfence_system();
    nbar += C;
    grid_barrier(a.bar, nbar, a.err);
    if (c == 0 && tid < a.nranks) {
      volatile unsigned long long* f = a.peerFlags[tid] + a.myRank;
      *f = a.epoch;
    }
    if (tid < a.nranks) {
      const volatile unsigned long long* f = a.myFlags + tid;
      const long long t0 = clock64();
      while (*f < a.epoch) {
        if (clock64() - t0 > (1ll << 34)) {
          if (a.err) *a.err = 1;
          break;
        }
      }
    }
    __syncthreads();
    __threadfence_system();
    bsrc = a.myWin + woff;
  }

  float x[RMAX], r[RMAX], u[RMAX], p[RMAX], s[RMAX], w[RMAX], idg[RMAX], wgt[RMAX];
  bool own[RMAX];
#pragma unroll
  for (int j = 0; j < RMAX; ++j) {
    const int lr = tid + j * kGThreads;
    own[j] = lr < nRows;
    const int g = row0 + lr;
    const float b = own[j] ? (bsrc ? __ldcg(bsrc + g) : a.rhsE[a.rowNode[g]]) : 0.f;
    idg[j] = own[j] ? a.invDiag[g] : 0.f;
    wgt[j] = own[j] ? (a.weight ? a.weight[g] : 1.f) : 0.f;
    x[j] = 0.f;
    r[j] = b;
    u[j] = idg[j] * b;
    p[j] = 0.f;
    s[j] = 0.f;
    w[j] = 0.f;
  }
  __syncthreads();  // matrix slice staged; `tot` / `wred` reused from here on

  double gamma = 0.0, gamma0 = 0.0, alpha = 0.0, beta = 0.0;
  int it = 0, par = 0;

  // one pass = publish u, SpMV, inner products, scalar recurrences
  auto spmv_dots = [&](bool first) {
    float* ub = a.uG + (size_t)par * a.NTpad;
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
      if (own[j]) ub[row0 + tid + j * kGThreads] = u[j];
    nbar += C;
    grid_barrier(a.bar, nbar, a.err);
    double pg = 0.0, pd = 0.0;
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
      if (own[j]) {
        const int lr = tid + j * kGThreads;
        float acc = 0.f;
        for (int k = 0; k < a.W; ++k)  // ascending ELL slot, as the other two paths
          acc += mv[(size_t)k * mstride + moff + lr] * __ldcg(ub + mc[(size_t)k * mstride + moff + lr]);
        w[j] = acc;
        const double ut = (double)u[j], wg = (double)wgt[j];
        pg += (double)r[j] * ut * wg;
        pd += (double)acc * ut * wg;
      }
    pg = warp_sum_dd(pg);
    pd = warp_sum_dd(pd);
    if (lane == 0) {
      wred[2 * wid] = pg;
      wred[2 * wid + 1] = pd;
    }
    __syncthreads();
    if (wid == 0) {
      double g2 = lane < kGThreads / 32 ? wred[2 * lane] : 0.0, d2 = lane < kGThreads / 32 ? wred[2 * lane + 1] : 0.0;
      g2 = warp_sum_dd(g2);
      d2 = warp_sum_dd(d2);
      if (lane == 0) {
        double* dst = a.redG + ((size_t)par * C + c) * 2;
        dst[0] = g2;
        dst[1] = d2;
      }
    }
    nbar += C;
    grid_barrier(a.bar, nbar, a.err);
    if (wid == 0) {
      // every CTA folds the same partials in the same order: lane-strided ascending sums, then the shuffle tree
      double g2 = 0.0, d2 = 0.0;
      for (int q = lane; q < C; q += 32) {
        g2 += __ldcg(a.redG + ((size_t)par * C + q) * 2);
        d2 += __ldcg(a.redG + ((size_t)par * C + q) * 2 + 1);
      }
      g2 = warp_sum_dd(g2);
      d2 = warp_sum_dd(d2);
      if (lane == 0) {
        tot[0] = g2;
        tot[1] = d2;
      }
    }
    __syncthreads();
    const double gn = tot[0], delta = tot[1];
    if (first) {
      gamma0 = gn;
      beta = 0.0;
      alpha = (delta > 0.0) ? gn / delta : 0.0;
    } else {
      const double b2 = (gamma > 0.0) ? gn / gamma : 0.0;
      const double den = (alpha != 0.0) ? delta - b2 * gn / alpha : delta;
      beta = b2;
      alpha = (den > 0.0) ? gn / den : 0.0;
    }
    gamma = gn;
    par ^= 1;
  };

  spmv_dots(true);
  for (it = 1; it <= a.maxIter; ++it) {
    const float al = (float)alpha, be = (float)beta;
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      const float pn = u[j] + be * p[j];
      const float sn = w[j] + be * s[j];
      p[j] = pn;
      s[j] = sn;
      x[j] = x[j] + al * pn;
      const float rn = r[j] - al * sn;
      r[j] = rn;
      u[j] = idg[j] * rn;
    }
    spmv_dots(false);
    if (it % a.checkEvery == 0 || it == a.maxIter)
      if (!(gamma > a.tol2 * gamma0)) break;
  }
  const int iters = min(it, a.maxIter);

  // solution -> L2 vector -> E-vector scatter (coarseLevel.cpp:216-221)
  {
    float* xb = a.uG + (size_t)par * a.NTpad;
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
      if (own[j]) xb[row0 + tid + j * kGThreads] = x[j];
    nbar += C;
    grid_barrier(a.bar, nbar, a.err);
    const long stride = (long)C * kGThreads;
    for (long n = (long)c * kGThreads + tid; n < a.Nlocal; n += stride) {
      const int t = a.tIndex[n];
      a.xE[n] = (t >= 0) ? __ldcg(xb + t) : 0.f;
    }
  }
  if (c == 0 && tid == 0) {
    a.S[0] = gamma;
    a.S[1] = gamma0;
    a.S[2] = (double)iters;
  }
}

template <int RMAX>
int launch_grid(const GridArgs& a, int C, size_t smem, cudaStream_t st, bool queryOnly)
{
  auto k = coarse_pcg_grid_kernel<RMAX>;
  cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  if (queryOnly) {
    int perSM = 0, dev = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k, kGThreads, smem) != cudaSuccess ||
        cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || perSM * sms < C) {
      cudaGetLastError();
      return 1;
    }
    return 0;
  }
  NRSB_CUDA(cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, 1, 1);
  cfg.blockDim = dim3(kGThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident, or the launch fails
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NRSB_CUDA(cudaLaunchKernelEx(&cfg, k, a));
  return NRSB_OK;
}

int dispatch_grid(const GridArgs& a, int C, int rmax, size_t smem, cudaStream_t st, bool queryOnly)
{
  switch (rmax) {
    case 1: return launch_grid<1>(a, C, smem, st, queryOnly);
    case 2: return launch_grid<2>(a, C, smem, st, queryOnly);
    case 4: return launch_grid<4>(a, C, smem, st, queryOnly);
    default: return 1;
  }
}

}  // namespace

// Picks the grid shape at setup for a system of `n` rows.  Leaves gridSize = 0 when it does not apply.
int coarseSolver_t::plan_grid()
{
  gridSize = 0;
  const int n = replicated ? NTg : NT;
  if ((multiRank && !replicated) || n <= 0) return NRSB_OK;
  int dev = 0, sms = 0, coop = 0;
  NRSB_CUDA(cudaGetDevice(&dev));
  NRSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  NRSB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return NRSB_OK;  // no co-residency guarantee: the cluster kernel or the multi-launch path take over
  int C = std::min(sms, (n + 63) / 64);
  const int RPC = ((n + C - 1) / C + 31) / 32 * 32;
  C = (n + RPC - 1) / RPC;
  int rmax = (RPC + kGThreads - 1) / kGThreads;
  if (rmax == 3) rmax = 4;
  if (rmax > 4) return NRSB_OK;
  const int W = replicated ? gEllWidth : ellWidth;
  const size_t fixed = (2 * (kGThreads / 32) + 2) * sizeof(double);
  const size_t mat = (size_t)W * RPC * 8;
  const int inSmem = fixed + mat <= 200 * 1024 ? 1 : 0;
  const size_t smem = fixed + (inSmem ? mat : 0);
  GridArgs a = {};
  if (dispatch_grid(a, C, rmax, smem, nullptr, true) != 0) return NRSB_OK;
  const int NTpad = (n + 3) / 4 * 4;
  int rc;
  if ((rc = gridU.alloc(2 * (size_t)NTpad))) return rc;
  if ((rc = gridRed.alloc(2 * (size_t)C * 2))) return rc;
  if ((rc = gridBar.alloc(1))) return rc;
  gridSize = C;
  gridRPC = RPC;
  gridRmax = rmax;
  gridSmem = smem;
  gridMatInSmem = inSmem;
  return NRSB_OK;
}

int coarseSolver_t::solve_grid(float* rhs, float* xE)
{
  elliptic_t* e = level->elliptic;
  GridArgs a = {};
  const int n = replicated ? NTg : NT;
  a.NT = n;
  a.W = replicated ? gEllWidth : ellWidth;
  a.RPC = gridRPC;
  a.NTpad = (n + 3) / 4 * 4;
  a.cols = replicated ? g_cols.p : d_cols.p;
  a.vals = replicated ? g_vals.p : d_vals.p;
  a.invDiag = replicated ? g_invDiag.p : invDiag.p;
  a.weight = replicated ? nullptr : d_weight.p;
  a.rowNode = d_rowNode.p;
  a.tIndex = replicated ? g_tIndex.p : d_tIndex.p;
  a.rhsE = rhs;
  a.xE = xE;
  a.Nlocal = e->mesh->Nlocal;
  a.maxIter = maxIter;
  a.checkEvery = checkEvery;
  a.tol2 = tol * tol;
  a.matInSmem = gridMatInSmem;
  a.S = scal.p + 8;
  a.uG = gridU.p;
  a.redG = gridRed.p;
  a.bar = gridBar.p;
  a.err = e->d_err;
  a.nranks = 1;
  if (replicated) {
    a.nranks = e->mesh->comm->nranks;
    a.myRank = e->mesh->comm->rank;
    a.nOwn = nOwn;
    a.ownG = d_ownG.p;
    a.ownNode = d_ownNode.p;
    a.peerWin = d_peerWin.p;
    a.peerFlags = d_peerWinFlags.p;
    a.myWin = rhsWindow;
    a.myFlags = (const unsigned long long*)(rhsWindow + 2 * (size_t)a.NTpad);
    a.epoch = ++winEpoch;
  }
  iterOnDevice = true;
  return dispatch_grid(a, gridSize, gridRmax, gridSmem, e->stream, false);
}

}  // namespace nrsb

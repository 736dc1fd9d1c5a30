// coarse_cluster.cu -- the whole coarse-grid PCG solve in ONE kernel on one thread-block cluster.
//
// The coarse problem of a rank is tiny (9 261 unknowns for 20^3 elements) and every Krylov iteration
// needs two global synchronisation points (the search direction must be visible before the SpMV, the
// inner products before the update).  As separate launches that is ~17 us per iteration of pure
// latency.  Here a cluster of up to 16 CTAs keeps everything on chip:
//   * the ELL matrix slice of each CTA and a full copy of the SpMV input vector live in shared memory;
//   * every thread owns up to RMAX rows; x, r, u, p, s, w of those rows stay in registers;
//   * the new u is broadcast to all CTAs with distributed-shared-memory stores, the inner-product
//     partials likewise; `cluster.sync()` (~0.2 us) replaces the kernel boundary;
//   * alpha/beta and the convergence test are evaluated redundantly by every thread (same bits).
// Same recurrences, same SpMV summation order and the same `checkEvery` convergence test as the
// multi-launch path in coarse.cu.
//
// Several ranks: the coarse problem is REPLICATED.  Every GPU holds the global assembled matrix; the
// kernel prologue all-gathers the right-hand side (each rank stores the entries it owns straight into
// every peer's window over NVLink, then raises an epoch flag), after which all GPUs run the identical
// solve on identical data and keep their own part of the solution.  One exchange per coarse solve
// instead of two per Krylov iteration.  When two copies of the vector no longer fit in shared memory
// the SpMV input goes through a global (L2-resident) buffer instead of distributed shared memory.
#include <cooperative_groups.h>

#include "host.hpp"

namespace cg = cooperative_groups;

namespace nrsb {

namespace {

constexpr int kCThreads = 1024;
constexpr int kMaxCluster = 16;

struct ClusterArgs {
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
  double* S;  // [0]=gamma, [1]=gamma0, [2]=iterations
  // SpMV input in global memory instead of distributed shared memory (large coarse grids)
  int uGlobal;
  float* uScratch;  // [2][NTpad]
  // replicated multi-rank solve: right-hand-side all-gather through peer windows
  int nranks, myRank;
  int nOwn;
  const int* ownG;     // global T index of the entries this rank owns
  const int* ownNode;  // E-vector node holding the value
  float* const* peerWin;                   // [nranks] windows: float[2][NTpad]
  unsigned long long* const* peerFlags;    // [nranks] flag arrays u64[nranks]
  const float* myWin;
  const unsigned long long* myFlags;
  unsigned long long epoch;
};

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int RMAX>
__global__ void __launch_bounds__(kCThreads, 1) coarse_pcg_cluster_kernel(const ClusterArgs a)
{
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int c = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  extern __shared__ __align__(16) unsigned char smraw[];
  double* red = reinterpret_cast<double*>(smraw);       // [2][kMaxCluster][2]  (written by peers)
  double* wred = red + 2 * kMaxCluster * 2;             // [32 warps][2]
  float* ubufS = reinterpret_cast<float*>(wred + 64);   // [2][NTpad]           (written by peers)
  float* mvals = ubufS + (a.uGlobal ? 0 : 2 * (size_t)a.NTpad);  // [W][RPC]
  float* ubuf = a.uGlobal ? a.uScratch : ubufS;
  int* mcols = reinterpret_cast<int*>(mvals + (size_t)a.W * a.RPC);

  const int row0 = c * a.RPC;
  const int nRows = max(0, min(a.RPC, a.NT - row0));

  // matrix slice -> shared memory (column-major: entry k of local row lr at [k*RPC + lr])
  const float* mv;
  const int* mc;
  int mstride, moff;
  if (a.matInSmem) {
    for (int k = 0; k < a.W; ++k)
      for (int lr = tid; lr < nRows; lr += kCThreads) {
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
    for (int i = c * kCThreads + tid; i < a.nOwn; i += C * kCThreads) {
      const float v = a.rhsE[a.ownNode[i]];
      const int g = a.ownG[i];
      for (int pr = 0; pr < a.nranks; ++pr) a.peerWin[pr][woff + g] = v;  // NVLink stores (own window included)
    }
    __threadfence_system();
    cluster.sync();
    if (c == 0 && tid < a.nranks) {
      volatile unsigned long long* f = a.peerFlags[tid] + a.myRank;
      *f = a.epoch;
    }
    if (tid < a.nranks) {
      const volatile unsigned long long* f = a.myFlags + tid;
      while (*f < a.epoch) {
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
    const int lr = tid + j * kCThreads;
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

  // make sure every CTA of the cluster is running before the first remote store
  cluster.sync();

  double gamma = 0.0, gamma0 = 0.0, alpha = 0.0, beta = 0.0;
  int it = 0, par = 0;

  // one pass = broadcast u, SpMV, inner products, scalar recurrences
  auto spmv_dots = [&](bool first) {
    float* ub = ubuf + (size_t)par * a.NTpad;
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
      if (own[j]) {
        const int g = row0 + tid + j * kCThreads;
        if (a.uGlobal) {
          ub[g] = u[j];
        } else {
          for (int q = 0; q < C; ++q) cluster.map_shared_rank(ub, q)[g] = u[j];
        }
      }
    cluster.sync();  // release/acquire at cluster scope: orders the global stores as well
    double pg = 0.0, pd = 0.0;
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
      if (own[j]) {
        const int lr = tid + j * kCThreads;
        float acc = 0.f;
        if (a.uGlobal) {
          for (int k = 0; k < a.W; ++k)
            acc += mv[(size_t)k * mstride + moff + lr] * __ldcg(ub + mc[(size_t)k * mstride + moff + lr]);
        } else {
          for (int k = 0; k < a.W; ++k)  // ascending column, as the multi-launch path
            acc += mv[(size_t)k * mstride + moff + lr] * ub[mc[(size_t)k * mstride + moff + lr]];
        }
        w[j] = acc;
        const double ut = (double)u[j], wg = (double)wgt[j];
        pg += (double)r[j] * ut * wg;
        pd += (double)acc * ut * wg;
      }
    pg = warp_sum_d(pg);
    pd = warp_sum_d(pd);
    if (lane == 0) {
      wred[2 * wid] = pg;
      wred[2 * wid + 1] = pd;
    }
    __syncthreads();
    if (wid == 0) {
      double g2 = wred[2 * lane], d2 = wred[2 * lane + 1];
      g2 = warp_sum_d(g2);
      d2 = warp_sum_d(d2);
      if (lane < C) {  // lane q delivers this CTA's partials to CTA q
        double* dst = cluster.map_shared_rank(red, lane) + ((size_t)par * kMaxCluster + c) * 2;
        dst[0] = g2;
        dst[1] = d2;
      }
    }
    cluster.sync();
    double gn = 0.0, delta = 0.0;
    for (int q = 0; q < C; ++q) {  // ascending CTA rank: same bits in every thread
      gn += red[((size_t)par * kMaxCluster + q) * 2];
      delta += red[((size_t)par * kMaxCluster + q) * 2 + 1];
    }
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

  // solution -> every CTA's shared memory -> E-vector scatter (coarseLevel.cpp:216-221)
  {
    float* xb = ubuf + (size_t)par * a.NTpad;
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
      if (own[j]) {
        const int g = row0 + tid + j * kCThreads;
        if (a.uGlobal) {
          xb[g] = x[j];
        } else {
          for (int q = 0; q < C; ++q) cluster.map_shared_rank(xb, q)[g] = x[j];
        }
      }
    cluster.sync();
    const long stride = (long)C * kCThreads;
    for (long n = (long)c * kCThreads + tid; n < a.Nlocal; n += stride) {
      const int t = a.tIndex[n];
      a.xE[n] = (t >= 0) ? (a.uGlobal ? __ldcg(xb + t) : xb[t]) : 0.f;
    }
  }
  if (c == 0 && tid == 0) {
    a.S[0] = gamma;
    a.S[1] = gamma0;
    a.S[2] = (double)iters;
  }
}

template <int RMAX>
int launch_cluster(const ClusterArgs& a, int C, size_t smem, cudaStream_t st, bool queryOnly)
{
  auto k = coarse_pcg_cluster_kernel<RMAX>;
  cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err == cudaSuccess && C > 8) err = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (err != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, 1, 1);
  cfg.blockDim = dim3(kCThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (queryOnly) {
    int n = 0;
    err = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    if (err != cudaSuccess || n < 1) {
      cudaGetLastError();
      return 1;
    }
    return 0;
  }
  NRSB_CUDA(cudaLaunchKernelEx(&cfg, k, a));
  return NRSB_OK;
}

int dispatch_cluster(const ClusterArgs& a, int C, int rmax, size_t smem, cudaStream_t st, bool queryOnly)
{
  switch (rmax) {
    case 1: return launch_cluster<1>(a, C, smem, st, queryOnly);
    case 2: return launch_cluster<2>(a, C, smem, st, queryOnly);
    case 4: return launch_cluster<4>(a, C, smem, st, queryOnly);
    case 8: return launch_cluster<8>(a, C, smem, st, queryOnly);
    default: return 1;
  }
}

}  // namespace

// Picks the cluster shape at setup for a system of `n` rows.  Leaves clusterSize = 0 when nothing fits.
int coarseSolver_t::plan_cluster()
{
  clusterSize = 0;
  const int n = replicated ? NTg : NT;
  if ((multiRank && !replicated) || n <= 0) return NRSB_OK;
  const size_t limit = 227 * 1024;
  const int NTpad = (n + 3) / 4 * 4;
  for (int C : {16, 8}) {
    const int RPC = ((n + C - 1) / C + 31) / 32 * 32;
    int rmax = (RPC + kCThreads - 1) / kCThreads;
    if (rmax == 3) rmax = 4;
    if (rmax > 4 && rmax <= 8) rmax = 8;
    if (rmax > 8) continue;
    const size_t fixed = (2 * kMaxCluster * 2 + 64) * sizeof(double);
    const size_t uBytes = 2 * (size_t)NTpad * sizeof(float);
    const size_t mat = (size_t)(replicated ? gEllWidth : ellWidth) * RPC * 8;
    // preference: everything in shared memory; then the matrix from L2; then the vector from L2 as well
    const int tryU[3] = {0, 0, 1}, tryM[3] = {1, 0, 0};
    int pick = -1;
    size_t smem = 0;
    for (int o = (variant == 2 ? 2 : 0); o < 3 && pick < 0; ++o) {  // variant 2 (tests): force the L2 path
      smem = fixed + (tryU[o] ? 0 : uBytes) + (tryM[o] ? mat : 0);
      if (smem <= limit) pick = o;
    }
    if (pick < 0) continue;
    ClusterArgs a = {};
    if (dispatch_cluster(a, C, rmax, smem, nullptr, true) != 0) continue;
    clusterSize = C;
    clusterRPC = RPC;
    clusterRmax = rmax;
    clusterSmem = smem;
    clusterMatInSmem = tryM[pick];
    clusterUGlobal = tryU[pick];
    if (clusterUGlobal) {
      int rc = uScratch.alloc(2 * (size_t)NTpad);
      if (rc) return rc;
    }
    break;
  }
  return NRSB_OK;
}

int coarseSolver_t::solve_cluster(float* rhs, float* xE)
{
  elliptic_t* e = level->elliptic;
  ClusterArgs a = {};
  const int n = replicated ? NTg : NT;
  a.NT = n;
  a.W = replicated ? gEllWidth : ellWidth;
  a.RPC = clusterRPC;
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
  a.matInSmem = clusterMatInSmem;
  a.S = scal.p + 8;
  a.uGlobal = clusterUGlobal;
  a.uScratch = uScratch.p;
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
  return dispatch_cluster(a, clusterSize, clusterRmax, clusterSmem, e->stream, false);
}

int coarseSolver_t::iterations()
{
  if (iterOnDevice) {
    double v = 0.0;
    if (cudaMemcpy(&v, scal.p + 10, sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess) lastIter = (int)v;
    iterOnDevice = false;
  }
  return lastIter;
}

}  // namespace nrsb

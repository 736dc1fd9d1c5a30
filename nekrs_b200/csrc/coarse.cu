// coarse.cu -- coarse-grid solve of the p-multigrid V-cycle.
//
// Takes the place of coarseLevel_t::solve (MG/coarseLevel.cpp:182-222: E->T gather, hypre BoomerAMG
// V-cycle on the CPU, T->E scatter).  hypre is third-party and outside this path;
// the same linear system is solved on the device instead:
//
//   * the N=1 operator is ASSEMBLED once (ellipticBuildFEMHex3D, MG/ellipticBuildFEM.cpp:73-305:
//     the GLL-quadrature element matrices summed over shared nodes) on this rank's unique unmasked
//     nodes ("T-vector", the rows of the masked ogs handle) as CSR in fp32;
//   * Jacobi-preconditioned CG in the Chronopoulos-Gear form (one reduction point per iteration),
//     TWO launches per iteration on one GPU: a fused vector update and a fused
//     SpMV + both inner products whose last block also updates alpha/beta on the device.  On several
//     GPUs the SpMV result of interface rows goes through the same NVLink halo exchange as every
//     other gather-scatter (oogs on the T-vector) and the inner products are all-reduced in-kernel.
//   * the host looks at one pair of scalars every `checkEvery` iterations.
//
// The coarse problem of a 20^3-element rank has 9 261 unknowns: every kernel is launch-latency
// bound, which is why the launch count (2 per iteration instead of 9) is what matters here.
#include <algorithm>
#include <cmath>
#include <map>

#include "host.hpp"
#include "reduce.cuh"

namespace nrsb {

namespace {

// scalar slots
enum { C_GAMMA = 0, C_DELTA, C_ALPHA, C_BETA, C_GAMMA0, C_TMP0, C_TMP1, C_COUNT = 16 };

struct CgPost {
  double* S;
  int first;
  __device__ __forceinline__ void operator()(double* tot) const
  {
    const double gn = tot[0], delta = tot[1];
    if (first) {
      S[C_GAMMA0] = gn;
      S[C_BETA] = 0.0;
      S[C_ALPHA] = (delta > 0.0) ? gn / delta : 0.0;
    } else {
      const double g = S[C_GAMMA], a = S[C_ALPHA];
      const double beta = (g > 0.0) ? gn / g : 0.0;
      const double den = (a != 0.0) ? delta - beta * gn / a : delta;
      S[C_BETA] = beta;
      S[C_ALPHA] = (den > 0.0) ? gn / den : 0.0;
    }
    S[C_GAMMA] = gn;
    S[C_DELTA] = delta;
  }
};

__global__ void __launch_bounds__(kBlockSize)
    coarse_init_kernel(int NT, const int* __restrict__ rowNode, const float* __restrict__ rhsE,
                       const float* __restrict__ invDiag, float* x, float* r, float* u, float* p, float* s)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NT) return;
  const float b = rhsE[rowNode[t]];
  x[t] = 0.f;
  r[t] = b;
  u[t] = invDiag[t] * b;
  p[t] = 0.f;
  s[t] = 0.f;
}

__global__ void __launch_bounds__(kBlockSize)
    coarse_update_kernel(int NT, const double* __restrict__ S, const float* __restrict__ invDiag,
                         const float* __restrict__ w, float* u, float* p, float* s, float* x, float* r)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NT) return;
  const float alpha = (float)S[C_ALPHA], beta = (float)S[C_BETA];
  const float pn = u[t] + beta * p[t];
  const float sn = w[t] + beta * s[t];
  p[t] = pn;
  s[t] = sn;
  x[t] = x[t] + alpha * pn;
  const float rn = r[t] - alpha * sn;
  r[t] = rn;
  u[t] = invDiag[t] * rn;
}

__device__ __forceinline__ float ell_row(int NT, int W, const int* __restrict__ cols, const float* __restrict__ vals,
                                         const float* __restrict__ u, long t)
{
  float acc = 0.f;
  for (int k = 0; k < W; k += 3) {  // W is a multiple of 3; summation order = ascending column, as CSR
    const int c0 = cols[(size_t)k * NT + t], c1 = cols[(size_t)(k + 1) * NT + t], c2 = cols[(size_t)(k + 2) * NT + t];
    const float v0 = vals[(size_t)k * NT + t], v1 = vals[(size_t)(k + 1) * NT + t], v2 = vals[(size_t)(k + 2) * NT + t];
    const float u0 = u[c0], u1 = u[c1], u2 = u[c2];
    acc += v0 * u0;
    acc += v1 * u1;
    acc += v2 * u2;
  }
  return acc;
}

__global__ void __launch_bounds__(kBlockSize)
    coarse_spmv_kernel(int NT, int W, const int* __restrict__ cols, const float* __restrict__ vals,
                       const float* __restrict__ u, float* __restrict__ w)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NT) return;
  w[t] = ell_row(NT, W, cols, vals, u, t);
}

__global__ void __launch_bounds__(kBlockSize)
    coarse_scatter_kernel(long Nlocal, const int* __restrict__ tIndex, const float* __restrict__ xT, float* xE)
{
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Nlocal) return;
  const int t = tIndex[n];
  xE[n] = (t >= 0) ? xT[t] : 0.f;
}

}  // namespace

int coarseSolver_t::setup(pMGLevel* lvl, int maxIter_, double tol_)
{
  level = lvl;
  maxIter = maxIter_;
  tol = tol_;
  elliptic_t* e = lvl->elliptic;
  mesh_t* mesh = e->mesh;
  ogs_t* ogs = e->ogs.get();
  const int Nq = mesh->Nq, Np = mesh->Np;
  const dlong E = mesh->Nelements;
  int rc;
  NRSB_REQUIRE(mesh->o_ggeo.p != nullptr, "coarse level needs fp64 geometric factors");

  // ---- T-vector numbering: rows of the masked gather-scatter handle (on-rank rows, then halo rows)
  const int nLoc = ogs->NlocalGather, nHalo = ogs->NhaloGather;
  NT = nLoc + nHalo;
  std::vector<int> tIndex(mesh->Nlocal, -1), rowNode(NT);
  std::vector<hlong> idsT(NT);
  for (int g = 0; g < nLoc; ++g) {
    for (dlong c = ogs->localGatherOffsets[g]; c < ogs->localGatherOffsets[g + 1]; ++c)
      tIndex[ogs->localGatherIds[c]] = g;
    rowNode[g] = ogs->localGatherIds[ogs->localGatherOffsets[g]];
  }
  for (int g = 0; g < nHalo; ++g) {
    for (dlong c = ogs->haloGatherOffsets[g]; c < ogs->haloGatherOffsets[g + 1]; ++c)
      tIndex[ogs->haloGatherIds[c]] = nLoc + g;
    rowNode[nLoc + g] = ogs->haloGatherIds[ogs->haloGatherOffsets[g]];
  }
  for (int t = 0; t < NT; ++t) idsT[t] = mesh->globalIds[rowNode[t]];

  // ---- element matrices (ellipticBuildFEMHex3D :141-225) and assembly
  std::vector<double> ggeo;
  if ((rc = mesh->o_ggeo.download(ggeo))) return rc;
  const std::vector<double>& D = mesh->D;
  const double lambda0 = e->lambda0Value, lambda1 = e->poisson ? 0.0 : e->lambda1Value;
  auto G = [&](dlong el, int c, int id) { return ggeo[(size_t)el * 7 * Np + (size_t)c * Np + id]; };
  std::vector<std::map<int, double>> rows(NT);
  for (dlong el = 0; el < E; ++el)
    for (int nz = 0; nz < Nq; nz++)
      for (int ny = 0; ny < Nq; ny++)
        for (int nx = 0; nx < Nq; nx++) {
          const int idn = nx + ny * Nq + nz * Nq * Nq;
          const int tn = tIndex[(size_t)el * Np + idn];
          if (tn < 0) continue;
          for (int mz = 0; mz < Nq; mz++)
            for (int my = 0; my < Nq; my++)
              for (int mx = 0; mx < Nq; mx++) {
                const int idm = mx + my * Nq + mz * Nq * Nq;
                const int tm = tIndex[(size_t)el * Np + idm];
                if (tm < 0) continue;
                double val = 0.;
                if (ny == my && nz == mz)
                  for (int k = 0; k < Nq; k++)
                    val += G(el, 0, k + ny * Nq + nz * Nq * Nq) * D[nx + k * Nq] * D[mx + k * Nq];
                if (nz == mz) {
                  val += G(el, 1, mx + ny * Nq + nz * Nq * Nq) * D[nx + mx * Nq] * D[my + ny * Nq];
                  val += G(el, 1, nx + my * Nq + nz * Nq * Nq) * D[mx + nx * Nq] * D[ny + my * Nq];
                }
                if (ny == my) {
                  val += G(el, 4, mx + ny * Nq + nz * Nq * Nq) * D[nx + mx * Nq] * D[mz + nz * Nq];
                  val += G(el, 4, nx + ny * Nq + mz * Nq * Nq) * D[mx + nx * Nq] * D[nz + mz * Nq];
                }
                if (nx == mx && nz == mz)
                  for (int k = 0; k < Nq; k++)
                    val += G(el, 2, nx + k * Nq + nz * Nq * Nq) * D[ny + k * Nq] * D[my + k * Nq];
                if (nx == mx) {
                  val += G(el, 3, nx + my * Nq + nz * Nq * Nq) * D[ny + my * Nq] * D[mz + nz * Nq];
                  val += G(el, 3, nx + ny * Nq + mz * Nq * Nq) * D[my + ny * Nq] * D[nz + mz * Nq];
                }
                if (nx == mx && ny == my)
                  for (int k = 0; k < Nq; k++)
                    val += G(el, 5, nx + ny * Nq + k * Nq * Nq) * D[nz + k * Nq] * D[mz + k * Nq];
                double valDiag = 0.;
                if (idn == idm) valDiag = G(el, 6, idn) * lambda1;
                rows[tn][tm] += lambda0 * val + lambda1 * valDiag;
              }
        }
  // ELL storage, column-major (entry k of row t at [k*NT + t]): fixed trip count -> the loads of one
  // row are independent and coalesced across rows.  Padding: column = own row, value = 0.
  ellWidth = 0;
  for (int t = 0; t < NT; ++t) ellWidth = std::max(ellWidth, (int)rows[t].size());
  ellWidth = (ellWidth + 2) / 3 * 3;
  std::vector<int> cols((size_t)ellWidth * NT);
  std::vector<float> vals((size_t)ellWidth * NT, 0.f), diag(NT, 0.f);
  for (int t = 0; t < NT; ++t) {
    int k = 0;
    for (auto& kv : rows[t]) {
      cols[(size_t)k * NT + t] = kv.first;
      vals[(size_t)k * NT + t] = (float)kv.second;
      if (kv.first == t) diag[t] = (float)kv.second;
      ++k;
    }
    for (; k < ellWidth; ++k) cols[(size_t)k * NT + t] = t;
  }
  if ((rc = d_cols.upload(cols))) return rc;
  if ((rc = d_vals.upload(vals))) return rc;
  if ((rc = d_rowNode.upload(rowNode))) return rc;
  if ((rc = d_tIndex.upload(tIndex))) return rc;

  // ---- T-vector exchange handle (interface rows hold partial sums after the SpMV)
  ogsT.reset(new ogs_t());
  if ((rc = ogsT->setup(NT, idsT.data(), mesh->topo.nranks > 1 ? &mesh->topo : nullptr))) return rc;
  oogsT.reset(new oogs_t());
  if ((rc = oogsT->setup(ogsT.get(), mesh->comm, 1))) return rc;
  multiRank = ogsT->NhaloGather > 0;
  // weights: every unique global node counts once in the inner products
  {
    std::vector<float> wt(NT);
    for (int t = 0; t < NT; ++t) wt[t] = (float)ogsT->invDegree[t];
    if ((rc = d_weight.upload(wt))) return rc;
  }
  // assembled diagonal (global sum on interface rows) -> Jacobi preconditioner
  {
    dbuf<float> d;
    if ((rc = d.upload(diag))) return rc;
    if ((rc = oogsT->startFinish<float>(d.p, 1, 0, gs_op::add, 0, nullptr, nullptr))) return rc;
    NRSB_CUDA(cudaDeviceSynchronize());
    if ((rc = d.download(diag))) return rc;
    for (auto& v : diag) v = 1.0f / v;
    if ((rc = invDiag.upload(diag))) return rc;
  }
  if ((rc = x.alloc(NT))) return rc;
  if ((rc = r.alloc(NT))) return rc;
  if ((rc = u.alloc(NT))) return rc;
  if ((rc = p.alloc(NT))) return rc;
  if ((rc = s.alloc(NT))) return rc;
  if ((rc = w.alloc(NT))) return rc;
  if ((rc = scal.alloc(C_COUNT))) return rc;
  if (multiRank && mesh->comm && mesh->comm->nranks > 1 && !e->options.compareArgs("COARSE SOLVER REPLICATED", "FALSE"))
    if ((rc = setup_replicated(idsT, rowNode, tIndex, rows))) return rc;
  if ((rc = plan_cluster())) return rc;
  return plan_grid();
}

coarseSolver_t::~coarseSolver_t()
{
  for (void* b : peerWinBase)
    if (b && b != (void*)rhsWindow) cudaIpcCloseMemHandle(b);
  if (rhsWindow) cudaFree(rhsWindow);
}

// Replicated coarse problem: global numbering of the unique unmasked coarse nodes, the global assembled
// matrix (every rank's element contributions summed in ascending rank order: identical bits everywhere),
// who pushes which right-hand-side entry, and the peer-mapped windows the entries are pushed into.
int coarseSolver_t::setup_replicated(const std::vector<hlong>& idsT, const std::vector<int>& rowNode,
                                     const std::vector<int>& tIndex, const std::vector<std::map<int, double>>& rows)
{
  comm_t* comm = level->elliptic->mesh->comm;
  const int nr = comm->nranks, me = comm->rank;
  int rc;
  // ---- global numbering
  std::vector<long> counts(nr, 0);
  counts[me] = NT;
  comm->allgather_bytes(counts.data(), sizeof(long));
  long maxNT = 1;
  for (long c : counts) maxNT = std::max(maxNT, c);
  std::vector<hlong> allIds((size_t)nr * maxNT, -1);
  std::copy(idsT.begin(), idsT.end(), allIds.begin() + (size_t)me * maxNT);
  comm->allgather_bytes(allIds.data(), (size_t)maxNT * sizeof(hlong));
  std::vector<hlong> G;
  for (int r = 0; r < nr; ++r) G.insert(G.end(), allIds.begin() + (size_t)r * maxNT, allIds.begin() + (size_t)r * maxNT + counts[r]);
  std::sort(G.begin(), G.end());
  G.erase(std::unique(G.begin(), G.end()), G.end());
  NTg = (int)G.size();
  auto gidx = [&](hlong id) { return (int)(std::lower_bound(G.begin(), G.end(), id) - G.begin()); };
  std::vector<int> owner(NTg, -1);
  for (int r = 0; r < nr; ++r)
    for (long i = 0; i < counts[r]; ++i) {
      const int g = gidx(allIds[(size_t)r * maxNT + i]);
      if (owner[g] < 0) owner[g] = r;
    }
  std::vector<int> gOfT(NT), ownG, ownNode;
  for (int t = 0; t < NT; ++t) {
    gOfT[t] = gidx(idsT[t]);
    if (owner[gOfT[t]] == me) {
      ownG.push_back(gOfT[t]);
      ownNode.push_back(rowNode[t]);
    }
  }
  nOwn = (int)ownG.size();
  // ---- global matrix: all-gather the local triplets
  std::vector<int> tr, tc;
  std::vector<double> tv;
  for (int t = 0; t < NT; ++t)
    for (auto& kv : rows[t]) {
      tr.push_back(gOfT[t]);
      tc.push_back(gOfT[kv.first]);
      tv.push_back(kv.second);
    }
  std::vector<long> nnz(nr, 0);
  nnz[me] = (long)tr.size();
  comm->allgather_bytes(nnz.data(), sizeof(long));
  long maxNnz = 1;
  for (long c : nnz) maxNnz = std::max(maxNnz, c);
  std::vector<int> allR((size_t)nr * maxNnz, 0), allC((size_t)nr * maxNnz, 0);
  std::vector<double> allV((size_t)nr * maxNnz, 0.0);
  std::copy(tr.begin(), tr.end(), allR.begin() + (size_t)me * maxNnz);
  std::copy(tc.begin(), tc.end(), allC.begin() + (size_t)me * maxNnz);
  std::copy(tv.begin(), tv.end(), allV.begin() + (size_t)me * maxNnz);
  comm->allgather_bytes(allR.data(), (size_t)maxNnz * sizeof(int));
  comm->allgather_bytes(allC.data(), (size_t)maxNnz * sizeof(int));
  comm->allgather_bytes(allV.data(), (size_t)maxNnz * sizeof(double));
  std::vector<std::map<int, double>> grow(NTg);
  for (int r = 0; r < nr; ++r)
    for (long i = 0; i < nnz[r]; ++i) grow[allR[(size_t)r * maxNnz + i]][allC[(size_t)r * maxNnz + i]] += allV[(size_t)r * maxNnz + i];
  gEllWidth = 0;
  for (int g = 0; g < NTg; ++g) gEllWidth = std::max(gEllWidth, (int)grow[g].size());
  gEllWidth = (gEllWidth + 2) / 3 * 3;
  std::vector<int> cols((size_t)gEllWidth * NTg);
  std::vector<float> vals((size_t)gEllWidth * NTg, 0.f), idg(NTg, 0.f);
  for (int g = 0; g < NTg; ++g) {
    int k = 0;
    for (auto& kv : grow[g]) {
      cols[(size_t)k * NTg + g] = kv.first;
      vals[(size_t)k * NTg + g] = (float)kv.second;
      if (kv.first == g) idg[g] = 1.0f / (float)kv.second;
      ++k;
    }
    for (; k < gEllWidth; ++k) cols[(size_t)k * NTg + g] = g;
  }
  std::vector<int> tIndexG(tIndex.size());
  for (size_t n = 0; n < tIndex.size(); ++n) tIndexG[n] = tIndex[n] >= 0 ? gOfT[tIndex[n]] : -1;
  if ((rc = g_cols.upload(cols))) return rc;
  if ((rc = g_vals.upload(vals))) return rc;
  if ((rc = g_invDiag.upload(idg))) return rc;
  if ((rc = g_tIndex.upload(tIndexG))) return rc;
  if ((rc = d_ownG.upload(ownG))) return rc;
  if ((rc = d_ownNode.upload(ownNode))) return rc;
  // ---- peer-mapped right-hand-side windows
  const size_t NTpad = ((size_t)NTg + 3) / 4 * 4;
  const size_t bytes = 2 * NTpad * sizeof(float) + (size_t)nr * sizeof(unsigned long long);
  NRSB_CUDA(cudaMalloc((void**)&rhsWindow, bytes));
  NRSB_CUDA(cudaMemset(rhsWindow, 0, bytes));
  NRSB_CUDA(cudaDeviceSynchronize());
  std::vector<cudaIpcMemHandle_t> handles(nr);
  NRSB_CUDA(cudaIpcGetMemHandle(&handles[me], rhsWindow));
  comm->allgather_bytes(handles.data(), sizeof(cudaIpcMemHandle_t));
  peerWinBase.assign(nr, nullptr);
  std::vector<float*> win(nr);
  std::vector<unsigned long long*> flg(nr);
  for (int p = 0; p < nr; ++p) {
    void* base = rhsWindow;
    if (p != me) NRSB_CUDA(cudaIpcOpenMemHandle(&base, handles[p], cudaIpcMemLazyEnablePeerAccess));
    peerWinBase[p] = base;
    win[p] = (float*)base;
    flg[p] = (unsigned long long*)((float*)base + 2 * NTpad);
  }
  if ((rc = d_peerWin.upload(win))) return rc;
  if ((rc = d_peerWinFlags.upload(flg))) return rc;
  comm->barrier();
  replicated = true;
  return NRSB_OK;
}

int coarseSolver_t::variant = 1;

// w = A u ; gamma = (r,u) ; delta = (w,u) ; alpha, beta updated by the last block
int coarseSolver_t::spmv_dots(bool first)
{
  elliptic_t* e = level->elliptic;
  cudaStream_t st = e->stream;
  double* S = scal.p;
  const int* cl = d_cols.p;
  const int W = ellWidth, nt = NT;
  const float* vl = d_vals.p;
  const float *up = u.p, *rp = r.p, *wt = d_weight.p;
  float* wp = w.p;
  CgPost post{S, first ? 1 : 0};
  if (!multiRank) {
    auto op = [=] __device__(long t, double* acc) {
      const float a = ell_row(nt, W, cl, vl, up, t);
      wp[t] = a;
      const double ut = (double)up[t], wgt = (double)wt[t];
      acc[0] += (double)rp[t] * ut * wgt;
      acc[1] += (double)a * ut * wgt;
    };
    return reduce_launch<2>(NT, op, 2, S + C_TMP0, e->ws, st, post, 1);
  }
  coarse_spmv_kernel<<<(NT + kBlockSize - 1) / kBlockSize, kBlockSize, 0, st>>>(NT, W, cl, vl, up, wp);
  NRSB_CHECK_LAUNCH();
  int rc = oogsT->startFinish<float>(wp, 1, 0, gs_op::add, 0, nullptr, st);
  if (rc) return rc;
  auto op = [=] __device__(long t, double* acc) {
    const double ut = (double)up[t], wgt = (double)wt[t];
    acc[0] += (double)rp[t] * ut * wgt;
    acc[1] += (double)wp[t] * ut * wgt;
  };
  return reduce_launch<2>(NT, op, 2, S + C_TMP0, e->ws, st, post, 1);
}

int coarseSolver_t::solve(float* rhs, float* xE)
{
  elliptic_t* e = level->elliptic;
  cudaStream_t st = e->stream;
  const int grid = (NT + kBlockSize - 1) / kBlockSize;
  double* S = scal.p;
  int rc;
  if (variant >= 1 && (!multiRank || replicated)) {
    const int n = replicated ? NTg : NT;
    const bool preferGrid = variant == 3 || (variant == 1 && (n > kGridRows || clusterSize == 0));
    if (gridSize > 0 && (preferGrid || clusterSize == 0) && variant != 2 && variant != 4) return solve_grid(rhs, xE);
    if (clusterSize > 0) return solve_cluster(rhs, xE);
  }
  iterOnDevice = false;
  lastIter = 0;
  if (NT > 0) {
    coarse_init_kernel<<<grid, kBlockSize, 0, st>>>(NT, d_rowNode.p, rhs, invDiag.p, x.p, r.p, u.p, p.p, s.p);
    NRSB_CHECK_LAUNCH();
  }
  if ((rc = spmv_dots(true))) return rc;
  int it = 0;
  for (it = 1; it <= maxIter; ++it) {
    if (NT > 0) {
      coarse_update_kernel<<<grid, kBlockSize, 0, st>>>(NT, S, invDiag.p, w.p, u.p, p.p, s.p, x.p, r.p);
      NRSB_CHECK_LAUNCH();
    }
    if ((rc = spmv_dots(false))) return rc;
    if (it % checkEvery == 0 || it == maxIter) {
      NRSB_CUDA(cudaMemcpyAsync(e->h_scal + 32, S + C_GAMMA, sizeof(double), cudaMemcpyDeviceToHost, st));
      NRSB_CUDA(cudaMemcpyAsync(e->h_scal + 33, S + C_GAMMA0, sizeof(double), cudaMemcpyDeviceToHost, st));
      NRSB_CUDA(cudaStreamSynchronize(st));
      if (!(e->h_scal[32] > tol * tol * e->h_scal[33])) break;
    }
  }
  lastIter = std::min(it, maxIter);
  const long Nlocal = e->mesh->Nlocal;
  if (Nlocal > 0) {
    coarse_scatter_kernel<<<(unsigned)((Nlocal + kBlockSize - 1) / kBlockSize), kBlockSize, 0, st>>>(
        Nlocal, d_tIndex.p, x.p, xE);
    NRSB_CHECK_LAUNCH();
  }
  return NRSB_OK;
}

}  // namespace nrsb

// elliptic.cpp -- host control flow of the elliptic solver.
//
// Restates, with identical semantics: ellipticSolveSetup (ellipticSetup.cpp:116-327), ellipticOgs
// (ellipticOgs.cpp:4-134), ellipticAx / ellipticOperator (ellipticOperator.cpp:31-172),
// ellipticApplyMask (ellipticApplyMask.cpp:3-27), ellipticZeroMean (ellipticZeroMean.cpp:31-44),
// ellipticSolve (ellipticSolve.cpp:32-190), ellipticPreconditioner (ellipticPreconditioner.cpp:33-84),
// pcg (PCG.cpp:33-203), pgmres (PGMRES.cpp:31-340), ellipticUpdateJacobi (ellipticUpdateJacobi.cpp:32-115).
//
// B200-first differences (results unchanged):
//  * Krylov scalars stay on the device: dot products are reduced (and all-reduced across GPUs) by
//    one kernel each and consumed by the next kernel through DevScalar; the host reads ONE value
//    per iteration (the residual norm, for the convergence test) instead of three blocking
//    device->host copies + three MPI_Allreduce (PCG.cpp:55-74, linAlg.cpp:1000-1025).
//  * `r -= alpha Ap`, `x += alpha p` and the weighted norm are a single kernel.
//  * mask + on-rank gather-scatter + halo unpack are a single kernel after Ax.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "host.hpp"
#include "halo.cuh"
#include "projection.hpp"

namespace nrsb {

// device scalar slots
enum { S_RDOTZ = 0, S_RDOTZ_OLD, S_PAP, S_RDOTR, S_ZDOTAP, S_ALPHA, S_BETA, S_SUM, S_NORM, S_DONE, S_ITERDONE, S_CURNORM,
       S_GMRES = 16, S_COUNT = 64 };

elliptic_t::elliptic_t() {}
elliptic_t::~elliptic_t()
{
  if (h_scal) cudaFreeHost(h_scal);
  if (h_err) cudaFreeHost(h_err);
}

int elliptic_t::read_scalars(int first, int count, double* out)
{
  NRSB_CUDA(cudaMemcpyAsync(h_scal + first, o_scal.p + first, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
  NRSB_CUDA(cudaStreamSynchronize(stream));
  for (int i = 0; i < count; ++i) out[i] = h_scal[first + i];
  if (comm && comm->peer_timeout()) {
    set_last_error("a device-side wait for a peer rank timed out (halo flags / all-reduce): a rank is missing");
    return NRSB_ERR_CUDA;
  }
  if (h_err && *h_err) {
    set_last_error("a device-wide wait inside a fused launch timed out: its CTAs were not co-resident (MPS / MIG / "
                   "fewer SMs than the launch assumes); results of that launch are invalid");
    return NRSB_ERR_CUDA;
  }
  return NRSB_OK;
}

// ------------------------------------------------------------------------------------------
// ellipticOgs: Dirichlet mask ids + masked gather-scatter handle
// ------------------------------------------------------------------------------------------
static void face_nodes(int N, std::vector<std::vector<int>>& fn)
{
  // meshBasisHex3D.cpp:53-82: f0 t=-1, f1 s=-1, f2 r=+1, f3 s=+1, f4 r=-1, f5 t=+1, ascending node index
  const int Nq = N + 1;
  fn.assign(6, {});
  for (int n = 0; n < Nq * Nq * Nq; ++n) {
    const int i = n % Nq, j = (n / Nq) % Nq, k = n / (Nq * Nq);
    if (k == 0) fn[0].push_back(n);
    if (j == 0) fn[1].push_back(n);
    if (i == N) fn[2].push_back(n);
    if (j == N) fn[3].push_back(n);
    if (i == 0) fn[4].push_back(n);
    if (k == N) fn[5].push_back(n);
  }
}

// ellipticOgs for a block solver (ellipticOgs.cpp:17-131 with nFields > 1; ellipticSetup.cpp:192-226): boundary flags
// per field (EToB[f + 6 e + fld * 6 E]), mask ids n + fld * fieldOffset, and the UNMASKED mesh numbering for all fields.
static int ellipticOgsBlock(mesh_t* mesh, const std::vector<int>& EToB, elliptic_t* elliptic)
{
  const int largeNumber = 1 << 20;
  const dlong Nlocal = mesh->Nlocal;
  const int Nfields = elliptic->Nfields;
  const dlong offset = elliptic->fieldOffset;
  NRSB_REQUIRE(EToB.size() >= (size_t)Nfields * mesh->Nelements * 6, "EToB needs Nfields * Nelements * 6 entries");
  std::vector<std::vector<int>> fn;
  face_nodes(mesh->N, fn);
  std::vector<char> isMasked((size_t)Nfields * Nlocal, 0);
  int rc;
  for (int fld = 0; fld < Nfields; ++fld) {
    std::vector<double> mapB(Nlocal, (double)largeNumber);
    for (dlong e = 0; e < mesh->Nelements; ++e)
      for (int f = 0; f < 6; ++f) {
        const int bc = EToB[f + (size_t)e * 6 + (size_t)fld * mesh->Nelements * 6];
        if (bc > 0)
          for (int n : fn[f]) {
            double& m = mapB[n + (size_t)e * mesh->Np];
            m = std::min((double)bc, m);
          }
      }
    dbuf<double> d;
    if ((rc = d.upload(mapB))) return rc;
    if ((rc = mesh->oogs->startFinish<double>(d.p, 1, 0, gs_op::min, 0, nullptr, nullptr))) return rc;
    NRSB_CUDA(cudaDeviceSynchronize());
    if ((rc = d.download(mapB))) return rc;
    for (dlong n = 0; n < Nlocal; ++n) isMasked[(size_t)fld * Nlocal + n] = mapB[n] == 1.0;  // DIRICHLET
  }
  elliptic->maskIds.clear();
  std::vector<dlong> loc, glo;
  for (int fld = 0; fld < Nfields; ++fld) {
    for (dlong n = 0; n < Nlocal; ++n)
      if (isMasked[(size_t)fld * Nlocal + n]) elliptic->maskIds.push_back(n + fld * offset);
    for (dlong e : mesh->localGatherElementList)
      for (int q = 0; q < mesh->Np; ++q)
        if (isMasked[(size_t)fld * Nlocal + (size_t)e * mesh->Np + q]) loc.push_back(e * mesh->Np + q + fld * offset);
    for (dlong e : mesh->globalGatherElementList)
      for (int q = 0; q < mesh->Np; ++q)
        if (isMasked[(size_t)fld * Nlocal + (size_t)e * mesh->Np + q]) glo.push_back(e * mesh->Np + q + fld * offset);
  }
  elliptic->Nmasked = (dlong)elliptic->maskIds.size();
  elliptic->NmaskedLocal = (dlong)loc.size();
  elliptic->NmaskedGlobal = (dlong)glo.size();
  if ((rc = elliptic->o_maskIds.upload(elliptic->maskIds))) return rc;
  if ((rc = elliptic->o_maskIdsLocal.upload(loc))) return rc;
  if ((rc = elliptic->o_maskIdsGlobal.upload(glo))) return rc;
  elliptic->ogs.reset(new ogs_t());
  if ((rc = elliptic->ogs->setup(Nlocal, mesh->globalIds.data(), mesh->topo.nranks > 1 ? &mesh->topo : nullptr)))
    return rc;
  elliptic->oogs.reset(new oogs_t());
  if ((rc = elliptic->oogs->setup(elliptic->ogs.get(), mesh->comm, Nfields))) return rc;
  elliptic->o_invDegree = elliptic->ogs->d_invDegree;
  elliptic->o_invDegreePfloat = elliptic->ogs->d_invDegreePfloat;
  // weights of the block inner products: invDegree per field, zero in the padding between Nlocal and fieldOffset
  std::vector<double> w((size_t)Nfields * offset, 0.0);
  for (int fld = 0; fld < Nfields; ++fld)
    for (dlong n = 0; n < Nlocal; ++n) w[(size_t)fld * offset + n] = elliptic->ogs->invDegree[n];
  return elliptic->o_weightBlock.upload(w);
}

int ellipticOgs(mesh_t* mesh, const std::vector<int>& EToB, elliptic_t* elliptic)
{
  if (elliptic->Nfields > 1) return ellipticOgsBlock(mesh, EToB, elliptic);
  const int largeNumber = 1 << 20;
  const dlong Nlocal = mesh->Nlocal;
  std::vector<std::vector<int>> fn;
  face_nodes(mesh->N, fn);
  // node-wise BC flag, DIRICHLET (lowest id) wins
  std::vector<double> mapB(Nlocal, (double)largeNumber);
  for (dlong e = 0; e < mesh->Nelements; ++e)
    for (int f = 0; f < 6; ++f) {
      const int bc = EToB[f + (size_t)e * 6];
      if (bc > 0)
        for (int n : fn[f]) {
          double& m = mapB[n + (size_t)e * mesh->Np];
          m = std::min((double)bc, m);
        }
    }
  // gs-min over the unmasked handle (flags are small integers: exact in fp64)
  {
    dbuf<double> d;
    int rc;
    if ((rc = d.upload(mapB))) return rc;
    if ((rc = mesh->oogs->startFinish<double>(d.p, 1, 0, gs_op::min, 0, nullptr, nullptr))) return rc;
    NRSB_CUDA(cudaDeviceSynchronize());
    if ((rc = d.download(mapB))) return rc;
  }
  elliptic->maskIds.clear();
  for (dlong n = 0; n < Nlocal; ++n)
    if (mapB[n] == 1.0) elliptic->maskIds.push_back(n);  // DIRICHLET == 1 (elliptic.h:46)
  elliptic->Nmasked = (dlong)elliptic->maskIds.size();
  std::vector<char> isMasked(Nlocal, 0);
  for (dlong n : elliptic->maskIds) isMasked[n] = 1;
  std::vector<dlong> loc, glo;
  for (dlong e : mesh->localGatherElementList)
    for (int q = 0; q < mesh->Np; ++q)
      if (isMasked[(size_t)e * mesh->Np + q]) loc.push_back(e * mesh->Np + q);
  for (dlong e : mesh->globalGatherElementList)
    for (int q = 0; q < mesh->Np; ++q)
      if (isMasked[(size_t)e * mesh->Np + q]) glo.push_back(e * mesh->Np + q);
  elliptic->NmaskedLocal = (dlong)loc.size();
  elliptic->NmaskedGlobal = (dlong)glo.size();
  int rc;
  if ((rc = elliptic->o_maskIds.upload(elliptic->maskIds))) return rc;
  if ((rc = elliptic->o_maskIdsLocal.upload(loc))) return rc;
  if ((rc = elliptic->o_maskIdsGlobal.upload(glo))) return rc;

  // masked version of the global numbering (ellipticOgs.cpp:126-131)
  std::vector<hlong> maskedGlobalIds(mesh->globalIds);
  for (dlong n : elliptic->maskIds) maskedGlobalIds[n] = 0;
  elliptic->ogs.reset(new ogs_t());
  if ((rc = elliptic->ogs->setup(Nlocal, maskedGlobalIds.data(), mesh->topo.nranks > 1 ? &mesh->topo : nullptr)))
    return rc;
  elliptic->oogs.reset(new oogs_t());
  if ((rc = elliptic->oogs->setup(elliptic->ogs.get(), mesh->comm, elliptic->Nfields))) return rc;
  elliptic->o_invDegree = elliptic->ogs->d_invDegree;
  elliptic->o_invDegreePfloat = elliptic->ogs->d_invDegreePfloat;
  return NRSB_OK;
}

// ------------------------------------------------------------------------------------------
template <typename T>
struct prec_traits;
template <>
struct prec_traits<double> {
  static const double* ggeo(mesh_t* m) { return m->o_ggeo.p; }
  static const double* D(mesh_t* m) { return m->D.data(); }
  static const double* lambda0(elliptic_t* e) { return e->lambdaField ? e->o_lambda0Field : e->o_lambda0.p; }
  static const double* lambda1(elliptic_t* e) { return e->lambdaField ? e->o_lambda1Field : e->o_lambda1.p; }
  static const double* invDegree(elliptic_t* e) { return e->o_invDegree; }
  static constexpr int idx = 0;
};
template <>
struct prec_traits<float> {
  static const float* ggeo(mesh_t* m) { return m->o_ggeoPfloat.p; }
  static const float* D(mesh_t* m) { return m->Dpfloat.data(); }
  static const float* lambda0(elliptic_t* e) { return e->lambdaField ? e->o_lambda0FieldPfloat.p : e->o_lambda0Pfloat.p; }
  static const float* lambda1(elliptic_t* e) { return e->lambdaField ? e->o_lambda1FieldPfloat.p : e->o_lambda1Pfloat.p; }
  static const float* invDegree(elliptic_t* e) { return e->o_invDegreePfloat; }
  static constexpr int idx = 1;
};

// block solver: ellipticBlockPartialAxCoeffHex3D / ellipticStressPartialAxCoeffHex3D (ellipticSetup.cpp:240-249)
static int ellipticAxBlock(elliptic_t* elliptic, dlong NelementsList, const dlong* o_elementList, const double* o_q,
                           double* o_Aq)
{
  mesh_t* mesh = elliptic->mesh;
  const double* l0 = elliptic->lambdaField ? elliptic->o_lambda0Field : elliptic->o_lambda0.p;
  const double* l1 = elliptic->lambdaField ? elliptic->o_lambda1Field : elliptic->o_lambda1.p;
  if (elliptic->stressForm) {
    NRSB_REQUIRE(mesh->o_vgeo.p != nullptr, "stress form needs mesh->o_vgeo");
    return ax_stress_launch<double>(mesh->Nq, NelementsList, elliptic->fieldOffset, elliptic->loffset, o_elementList,
                                    mesh->o_vgeo.p, mesh->D.data(), l0, l1, elliptic->lambdaField ? 1 : 0, o_q, o_Aq,
                                    elliptic->stream);
  }
  NRSB_REQUIRE(mesh->o_ggeo.p != nullptr, "fp64 geometric factors are not resident");
  return ax_block_launch<double>(mesh->Nq, 1, NelementsList, elliptic->Nfields, elliptic->fieldOffset,
                                 elliptic->loffset, o_elementList, mesh->o_ggeo.p, mesh->D.data(), l0, l1,
                                 elliptic->lambdaField ? 1 : 0, o_q, o_Aq, elliptic->stream);
}
template <typename T>
static int ellipticAxBlockT(elliptic_t* elliptic, dlong n, const dlong* list, const T* o_q, T* o_Aq)
{
  if constexpr (sizeof(T) == 8) return ellipticAxBlock(elliptic, n, list, o_q, o_Aq);
  set_last_error("block solves run in fp64");
  return NRSB_ERR_INVALID;
}

template <typename T>
static int ellipticAxDot(elliptic_t* elliptic, dlong NelementsList, const dlong* o_elementList, const T* o_q, T* o_Aq,
                         AxDot* dot)
{
  if (dot) dot->n = 0;
  if (NelementsList == 0) return NRSB_OK;
  if (elliptic->Nfields > 1) return ellipticAxBlockT<T>(elliptic, NelementsList, o_elementList, o_q, o_Aq);
  mesh_t* mesh = elliptic->mesh;
  using P = prec_traits<T>;
  NRSB_REQUIRE(P::ggeo(mesh) != nullptr, "geometric factors of the requested precision are not resident");
  int variant = elliptic->ax_variant[P::idx];
  if (variant < 0) variant = ax_default_variant(mesh->Nq, (int)sizeof(T));
  return ax_launch<T>(mesh->Nq, variant, NelementsList, elliptic->loffset, o_elementList, P::ggeo(mesh), P::D(mesh),
                      P::lambda0(elliptic), P::lambda1(elliptic), elliptic->poisson ? 1 : 0, elliptic->lambdaField ? 1 : 0,
                      o_q, o_Aq, elliptic->stream, dot);
}

template <typename T>
int ellipticAx(elliptic_t* elliptic, dlong NelementsList, const dlong* o_elementList, const T* o_q, T* o_Aq)
{
  if (NelementsList == 0) return NRSB_OK;
  if (elliptic->Nfields > 1) return ellipticAxBlockT<T>(elliptic, NelementsList, o_elementList, o_q, o_Aq);
  mesh_t* mesh = elliptic->mesh;
  using P = prec_traits<T>;
  NRSB_REQUIRE(P::ggeo(mesh) != nullptr, "geometric factors of the requested precision are not resident");
  int variant = elliptic->ax_variant[P::idx];
  if (variant < 0) variant = ax_default_variant(mesh->Nq, (int)sizeof(T));
  return ax_launch<T>(mesh->Nq, variant, NelementsList, elliptic->loffset, o_elementList, P::ggeo(mesh), P::D(mesh),
                      P::lambda0(elliptic), P::lambda1(elliptic), elliptic->poisson ? 1 : 0, elliptic->lambdaField ? 1 : 0,
                      o_q, o_Aq, elliptic->stream);
}
template int ellipticAx<double>(elliptic_t*, dlong, const dlong*, const double*, double*);
template int ellipticAx<float>(elliptic_t*, dlong, const dlong*, const float*, float*);

template <typename T>
int ellipticApplyMask(elliptic_t* elliptic, T* o_x)
{
  return mask_launch<T>(elliptic->Nmasked, elliptic->o_maskIds.p, o_x, elliptic->stream);
}
template int ellipticApplyMask<double>(elliptic_t*, double*);
template int ellipticApplyMask<float>(elliptic_t*, float*);

// ellipticOperator (ellipticOperator.cpp:117-172).  With `overlap` the halo elements are done first,
// their partial sums are pushed to the neighbours (oogs::start), the interior elements follow while
// the NVLink stores are in flight, and oogs::finish folds everything (and the mask) in one kernel.
template <typename T>
int ellipticOperator(elliptic_t* elliptic, const T* o_q, T* o_Aq, bool masked, AxDot* dot)
{
  if (dot) dot->n = 0;
  mesh_t* mesh = elliptic->mesh;
  oogs_t* oogs = elliptic->oogs.get();
  int rc;
  const dlong nm = masked ? elliptic->Nmasked : 0;
  if (elliptic->Nfields > 1) {
    // block solver: the handle carries the UNMASKED numbering (ellipticOgs.cpp:121-131 builds a masked handle for one
    // field only), so masked nodes belong to rows: zero them first, then sum (all copies of a Dirichlet node are masked)
    if ((rc = ellipticAx<T>(elliptic, mesh->Nelements, mesh->o_elementList.p, o_q, o_Aq))) return rc;
    if (nm)
      if ((rc = mask_launch<T>(nm, elliptic->o_maskIds.p, o_Aq, elliptic->stream))) return rc;
    return oogs->startFinish<T>(o_Aq, elliptic->Nfields, elliptic->fieldOffset, gs_op::add, 0, nullptr, elliptic->stream);
  }
  using P = prec_traits<T>;
  const int axv = elliptic->ax_variant[P::idx] < 0 ? ax_default_variant(mesh->Nq, (int)sizeof(T))
                                                   : elliptic->ax_variant[P::idx];
  const bool gsInLaunch = elliptic->fusedGsAx && elliptic->poisson && !elliptic->lambdaField && mesh->Nq == 8 &&
                          elliptic->Nfields == 1 && axv >= 4;
  FusedRows FR;
  if (gsInLaunch) {
    if (!elliptic->fusedArrive.p) {
      if ((rc = elliptic->fusedArrive.alloc(1))) return rc;  // zero-initialised
      NRSB_CUDA(cudaDeviceSynchronize());
    }
    FR.R = oogs->ogs->rows;
    FR.R.nMasked = nm;
    FR.R.maskIds = elliptic->o_maskIds.p;
    FR.arrive = elliptic->fusedArrive.p;
    FR.err = elliptic->d_err;
    FR.target = elliptic->fusedArriveTarget;
  }
  if (gsInLaunch && oogs->ogs->NhaloGather == 0) {
    // single rank: the whole operator (Ax, mask, gather-scatter) is ONE launch
    rc = ax_tma_gs_launch<T>(mesh->Nq, mesh->Nelements, mesh->o_elementList.p, P::ggeo(mesh), P::D(mesh),
                             P::lambda0(elliptic), P::lambda1(elliptic), elliptic->poisson ? 1 : 0, o_q, o_Aq, nullptr,
                             &FR, elliptic->stream, dot);
    elliptic->fusedArriveTarget = FR.target;
    return rc;
  }
  // (the fused launches use the 6-stage ring, which has no room for the Helmholtz GwJ plane: Poisson only)
  if (elliptic->overlap && elliptic->fusedHaloAx && elliptic->poisson && !elliptic->lambdaField && mesh->Nq == 8 &&
      elliptic->Nfields == 1 &&
      elliptic->ax_variant[P::idx] != 0 && mesh->NglobalGatherElements > 0 && oogs->peers.size() <= 32) {
    // ONE launch: Ax over [halo elements, interior elements]; the last F.nPush CTAs of the grid do no element work,
    // they push the halo partial sums over NVLink as soon as the halo elements are stored (oogs::begin_fused sizes
    // nPush from the send table).  The mask moves to finish (masked nodes belong to no row of the masked handle,
    // so it commutes with every sum).
    FusedHalo F;
    if ((rc = oogs->begin_fused(&F, mesh->NglobalGatherElements, elliptic->fieldOffset))) return rc;
    F.err = elliptic->d_err;
    // developer aid (NRSB_OP_TIMING=1): per-kernel times of the pipelined operator, no host sync per step
    static const bool timing = getenv("NRSB_OP_TIMING") != nullptr;
    static std::vector<cudaEvent_t> ev;
    static int nrec = 0;
    constexpr int kRing = 50;
    if (timing) {
      if (ev.empty()) {
        ev.resize(3 * kRing);
        for (auto& x : ev) cudaEventCreate(&x);
      }
      cudaEventRecord(ev[3 * nrec], elliptic->stream);
      static unsigned long long* d_stamps = nullptr;
      if (!d_stamps) {
        cudaMalloc((void**)&d_stamps, 16 * sizeof(unsigned long long));
        cudaMemset(d_stamps, 0, 16 * sizeof(unsigned long long));
      }
      F.stamps = d_stamps;
    }
    if ((rc = ax_tma_fused_launch<T>(mesh->Nq, 5, mesh->Nelements, mesh->o_haloFirstElementList.p, P::ggeo(mesh),
                                     P::D(mesh), P::lambda0(elliptic), P::lambda1(elliptic), elliptic->poisson ? 1 : 0,
                                     o_q, o_Aq, F, elliptic->stream, dot)))
      return rc;
    if (timing) cudaEventRecord(ev[3 * nrec + 1], elliptic->stream);
    rc = oogs->finish<T>(o_Aq, 1, elliptic->fieldOffset, gs_op::add, nm, elliptic->o_maskIds.p, elliptic->stream);
    if (timing) {
      cudaEventRecord(ev[3 * nrec + 2], elliptic->stream);
      if (++nrec == kRing) {
        cudaEventSynchronize(ev[3 * kRing - 1]);
        double sa = 0, sb = 0;
        for (int i = 0; i < kRing; ++i) {
          float x = 0, y = 0;
          cudaEventElapsedTime(&x, ev[3 * i], ev[3 * i + 1]);
          cudaEventElapsedTime(&y, ev[3 * i + 1], ev[3 * i + 2]);
          sa += x;
          sb += y;
        }
        fprintf(stderr, "[rank %d] fused operator: Ax+push %.2f us, finish %.2f us (mean of %d, pipelined)\n",
                mesh->comm ? mesh->comm->rank : 0, sa / kRing * 1e3, sb / kRing * 1e3, kRing);
        unsigned long long h[16];
        cudaMemcpy(h, F.stamps, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr,
                "[rank %d]   last launch, ns after pusher start: halo elements stored %lld, values loaded %lld, pushed %lld, fenced %lld, "
                "flags %lld, axhelm CTA 0 done %lld, last axhelm CTA done %lld\n",
                mesh->comm ? mesh->comm->rank : 0, (long long)(h[1] - h[0]), (long long)(h[5] - h[0]), (long long)(h[2] - h[0]),
                (long long)(h[3] - h[0]), (long long)(h[4] - h[0]), (long long)(h[8] - h[0]), (long long)(h[9] - h[0]));
        nrec = 0;
      }
    }
    return rc;
  }
  // The reference's split (Ax on the halo elements, oogs::start, Ax on the interior, oogs::finish: four launches and a
  // system-scope fence here) only pays when the exchange is slow.  With the one-launch flag-in-data exchange
  // (oogs_t::exchange_ll) the unsplit form below -- Ax on all elements, then ONE gather-scatter + exchange launch --
  // is faster on every multigrid level; ENABLE GS COMM OVERLAP = SPLIT keeps the reference's sequence.
  if (elliptic->overlap && elliptic->splitOverlap) {
    if ((rc = ellipticAx<T>(elliptic, mesh->NglobalGatherElements, mesh->o_globalGatherElementList.p, o_q, o_Aq)))
      return rc;
    if (masked && elliptic->NmaskedGlobal)
      if ((rc = mask_launch<T>(elliptic->NmaskedGlobal, elliptic->o_maskIdsGlobal.p, o_Aq, elliptic->stream)))
        return rc;
    if ((rc = oogs->start<T>(o_Aq, elliptic->Nfields, elliptic->fieldOffset, gs_op::add, elliptic->stream))) return rc;
    if ((rc = ellipticAx<T>(elliptic, mesh->NlocalGatherElements, mesh->o_localGatherElementList.p, o_q, o_Aq)))
      return rc;
    return oogs->finish<T>(o_Aq, elliptic->Nfields, elliptic->fieldOffset, gs_op::add,
                           masked ? elliptic->NmaskedLocal : 0, elliptic->o_maskIdsLocal.p, elliptic->stream);
  }
  if ((rc = ellipticAxDot<T>(elliptic, mesh->Nelements, mesh->o_elementList.p, o_q, o_Aq, dot))) return rc;
  // (masked nodes carry id 0 in the masked handle: they belong to no row, on-rank or halo, so zeroing them commutes
  // with every sum and rides along in the gather-scatter launch)
  return oogs->startFinish<T>(o_Aq, elliptic->Nfields, elliptic->fieldOffset, gs_op::add, nm, elliptic->o_maskIds.p,
                              elliptic->stream);
}
template int ellipticOperator<double>(elliptic_t*, const double*, double*, bool, AxDot*);
template int ellipticOperator<float>(elliptic_t*, const float*, float*, bool, AxDot*);

// "testing Ax overlap" of ellipticSetup.cpp:255-302 (solver handle, fp64) and ellipticMultiGridSetup.cpp:123-150
// (levels, fp32): one warm-up, barrier, 10 operator applications, the max over ranks decides -- so every rank takes
// the same branch.  The default (TRUE) does not measure: the unsplit form with the one-launch exchange won on every
// level and GPU count measured in round 2.
template <typename T>
static int time_operator(elliptic_t* elliptic, T* q, T* Aq, double* seconds)
{
  comm_t* comm = elliptic->mesh->comm;
  const int Nsamples = 10;
  int rc;
  if ((rc = ellipticOperator<T>(elliptic, q, Aq, true, nullptr))) return rc;
  NRSB_CUDA(cudaStreamSynchronize(elliptic->stream));
  if (comm && comm->barrier) comm->barrier();
  cudaEvent_t a, b;
  NRSB_CUDA(cudaEventCreate(&a));
  NRSB_CUDA(cudaEventCreate(&b));
  NRSB_CUDA(cudaEventRecord(a, elliptic->stream));
  for (int test = 0; test < Nsamples && !rc; ++test) rc = ellipticOperator<T>(elliptic, q, Aq, true, nullptr);
  NRSB_CUDA(cudaEventRecord(b, elliptic->stream));
  NRSB_CUDA(cudaEventSynchronize(b));
  float ms = 0.f;
  NRSB_CUDA(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  if (rc) return rc;
  double t = (double)ms * 1e-3 / Nsamples;
  if (comm && comm->nranks > 1 && comm->allgather_bytes) {
    std::vector<double> all(comm->nranks, 0.0);
    all[comm->rank] = t;
    comm->allgather_bytes(all.data(), sizeof(double));
    for (double v : all) t = std::max(t, v);
  }
  *seconds = t;
  return NRSB_OK;
}

int ellipticChooseOverlap(elliptic_t* elliptic, int precision)
{
  options_t& options = elliptic->options;
  elliptic->splitOverlap = false;
  if (!elliptic->overlap) return NRSB_OK;
  if (options.compareArgs("ENABLE GS COMM OVERLAP", "UNSPLIT")) return NRSB_OK;
  if (options.compareArgs("ENABLE GS COMM OVERLAP", "SPLIT")) {
    elliptic->splitOverlap = true;
    return NRSB_OK;
  }
  if (!options.compareArgs("ENABLE GS COMM OVERLAP", "TIMED")) return NRSB_OK;
  if (elliptic->lambdaField && !elliptic->o_lambda0Field) return NRSB_OK;  // coefficient fields arrive after setup
  const size_t n = (size_t)elliptic->fieldOffset * elliptic->Nfields;
  int rc;
  for (int split = 0; split < 2; ++split) {
    elliptic->splitOverlap = split != 0;
    if (precision == 8) {
      dbuf<double> q, Aq;
      if ((rc = q.alloc(n)) || (rc = Aq.alloc(n))) return rc;
      if ((rc = time_operator<double>(elliptic, q.p, Aq.p, &elliptic->overlapTimes[split]))) return rc;
    } else {
      dbuf<float> q, Aq;
      if ((rc = q.alloc(n)) || (rc = Aq.alloc(n))) return rc;
      if ((rc = time_operator<float>(elliptic, q.p, Aq.p, &elliptic->overlapTimes[split]))) return rc;
    }
  }
  elliptic->splitOverlap = elliptic->overlapTimes[1] < elliptic->overlapTimes[0];
  if ((!elliptic->mesh->comm || elliptic->mesh->comm->rank == 0) && getenv("NRSB_VERBOSE"))
    fprintf(stderr, "testing Ax overlap (N=%d, %s): unsplit %.2es split %.2es (%s)\n", elliptic->mesh->N,
            precision == 8 ? "fp64" : "fp32", elliptic->overlapTimes[0], elliptic->overlapTimes[1],
            elliptic->splitOverlap ? "split" : "unsplit");
  return NRSB_OK;
}

int ellipticZeroMean(elliptic_t* elliptic, double* o_q)
{
  mesh_t* mesh = elliptic->mesh;
  const double Nglobal = (double)elliptic->mesh->NelementsGlobal * mesh->Np;
  int rc;
  if ((rc = sum_launch<double>(mesh->Nlocal, o_q, elliptic->o_scal.p + S_SUM, elliptic->ws, elliptic->stream)))
    return rc;
  DevScalar a = DevScalar::ratio(elliptic->o_scal.p + S_SUM, nullptr, -1.0 / Nglobal);
  return add_scalar_launch<double>(mesh->Nlocal, a, o_q, elliptic->stream);
}

// ellipticUpdateJacobi(elliptic, o_invDiagA) (ellipticUpdateJacobi.cpp:32-85): element diagonals on the device
// (diag.cu = ellipticBlockBuildDiagonalHex3D), gather-scatter, adyMany / padyMany.
template <typename T>
int ellipticBuildDiagonal(elliptic_t* elliptic, T* o_invDiagA)
{
  mesh_t* mesh = elliptic->mesh;
  using P = prec_traits<T>;
  int rc;
  const T* ggeo = P::ggeo(mesh);
  NRSB_REQUIRE(ggeo != nullptr, "geometric factors of the requested precision are not resident");
  if ((rc = build_diagonal_launch<T>(mesh->Nq, mesh->Nelements, elliptic->Nfields, elliptic->fieldOffset,
                                     elliptic->loffset, ggeo, P::D(mesh), P::lambda0(elliptic), P::lambda1(elliptic),
                                     elliptic->poisson ? 1 : 0, elliptic->lambdaField ? 1 : 0, o_invDiagA,
                                     elliptic->stream)))
    return rc;
  if ((rc = elliptic->oogs->startFinish<T>(o_invDiagA, elliptic->Nfields, elliptic->fieldOffset, gs_op::add, 0, nullptr,
                                           elliptic->stream)))
    return rc;
  return ady_many_launch<T>(mesh->Nlocal, elliptic->Nfields, elliptic->fieldOffset, T(1), o_invDiagA, elliptic->stream);
}
template int ellipticBuildDiagonal<double>(elliptic_t*, double*);
template int ellipticBuildDiagonal<float>(elliptic_t*, float*);

// ellipticUpdateJacobi(ellipticBase) (ellipticUpdateJacobi.cpp:87-115): refresh every inverse diagonal that depends
// on the coefficients -- the smoother diagonals of the multigrid levels (DAMPEDJACOBI) or the Jacobi preconditioner.
int ellipticUpdateJacobi(elliptic_t* ellipticBase)
{
  options_t& options = ellipticBase->options;
  precon_t* precon = ellipticBase->precon.get();
  if (!precon) return NRSB_OK;
  int rc;
  if (options.compareArgs("PRECONDITIONER", "MULTIGRID") && options.compareArgs("MULTIGRID SMOOTHER", "DAMPEDJACOBI")) {
    auto& levels = precon->MGSolver->levels;
    for (size_t k = 0; k < levels.size(); ++k) {
      pMGLevel* L = levels[k].get();
      const bool coarsest = (k + 1 == levels.size());
      if (coarsest && options.compareArgs("MULTIGRID COARSE SOLVE", "TRUE")) continue;
      if (!L->o_invDiagA.p) continue;
      if ((rc = ellipticBuildDiagonal<float>(L->elliptic, L->o_invDiagA.p))) return rc;
    }
  } else if (options.compareArgs("PRECONDITIONER", "JACOBI")) {
    if ((rc = ellipticBuildDiagonal<double>(ellipticBase, precon->o_invDiagA.p))) return rc;
  }
  return NRSB_OK;
}

// ellipticMultiGridUpdateLambda (MG/ellipticMultiGridUpdateLambda.cpp): level 0 gets the fp32 cast of the solver's
// coefficient fields, every coarser level the nodal interpolation (coarsen kernel with the fine->coarse
// interpolation matrix, ellipticBuildMultigridLevel.cpp:133-146) of the level above.
int ellipticMultiGridUpdateLambda(elliptic_t* elliptic)
{
  if (!elliptic->lambdaField) return NRSB_OK;
  mesh_t* mesh = elliptic->mesh;
  cudaStream_t st = elliptic->stream;
  int rc;
  // the solver's own fp32 copy (fp32 operator on the solver mesh)
  if (elliptic->o_lambda0FieldPfloat.n < (size_t)mesh->Nlocal)
    if ((rc = elliptic->o_lambda0FieldPfloat.alloc(mesh->Nlocal))) return rc;
  if ((rc = copy_d2f_launch(mesh->Nlocal, elliptic->o_lambda0Field, elliptic->o_lambda0FieldPfloat.p, st))) return rc;
  if (!elliptic->poisson) {
    if (elliptic->o_lambda1FieldPfloat.n < (size_t)mesh->Nlocal)
      if ((rc = elliptic->o_lambda1FieldPfloat.alloc(mesh->Nlocal))) return rc;
    if ((rc = copy_d2f_launch(mesh->Nlocal, elliptic->o_lambda1Field, elliptic->o_lambda1FieldPfloat.p, st))) return rc;
  }
  if (!elliptic->precon || !elliptic->precon->MGSolver) return NRSB_OK;
  auto& levels = elliptic->precon->MGSolver->levels;
  for (size_t k = 0; k < levels.size(); ++k) {
    elliptic_t* e = levels[k]->elliptic;
    const dlong Nl = e->mesh->Nlocal;
    if (e->o_lambda0FieldPfloat.n < (size_t)Nl)
      if ((rc = e->o_lambda0FieldPfloat.alloc(Nl))) return rc;
    if (!elliptic->poisson && e->o_lambda1FieldPfloat.n < (size_t)Nl)
      if ((rc = e->o_lambda1FieldPfloat.alloc(Nl))) return rc;
    if (k == 0) {
      if ((rc = copy_d2f_launch(Nl, elliptic->o_lambda0Field, e->o_lambda0FieldPfloat.p, st))) return rc;
      if (!elliptic->poisson)
        if ((rc = copy_d2f_launch(Nl, elliptic->o_lambda1Field, e->o_lambda1FieldPfloat.p, st))) return rc;
    } else {
      elliptic_t* f = levels[k - 1]->elliptic;
      NRSB_REQUIRE(!e->interpFromFine.empty(), "multigrid level has no fine-to-coarse interpolation matrix");
      if ((rc = transfer_dispatch(true, f->mesh->Nq, e->mesh->Nq, e->mesh->Nelements, e->interpFromFine.data(),
                                  f->o_lambda0FieldPfloat.p, e->o_lambda0FieldPfloat.p, st)))
        return rc;
      if (!elliptic->poisson)
        if ((rc = transfer_dispatch(true, f->mesh->Nq, e->mesh->Nq, e->mesh->Nelements, e->interpFromFine.data(),
                                    f->o_lambda1FieldPfloat.p, e->o_lambda1FieldPfloat.p, st)))
          return rc;
    }
    e->lambdaField = true;
  }
  return NRSB_OK;
}

// ELLIPTIC COEFF FIELD: switch the handle (and its multigrid levels) to per-node coefficients owned by the caller.
// nullptr, nullptr switches back to the constant coefficients of the setup.
int ellipticSetCoeffField(elliptic_t* elliptic, const double* o_lambda0, const double* o_lambda1)
{
  if (!o_lambda0) {
    elliptic->lambdaField = false;
    elliptic->o_lambda0Field = elliptic->o_lambda1Field = nullptr;
    if (elliptic->precon && elliptic->precon->MGSolver)
      for (auto& L : elliptic->precon->MGSolver->levels) L->elliptic->lambdaField = false;
    return ellipticUpdateJacobi(elliptic);
  }
  NRSB_REQUIRE(elliptic->poisson || o_lambda1, "lambda1 field is NULL for a Helmholtz operator");
  elliptic->lambdaField = true;
  elliptic->o_lambda0Field = o_lambda0;
  elliptic->o_lambda1Field = o_lambda1;
  int rc;
  if ((rc = ellipticMultiGridUpdateLambda(elliptic))) return rc;
  return ellipticUpdateJacobi(elliptic);
}

// ------------------------------------------------------------------------------------------
// setup
// ------------------------------------------------------------------------------------------
int elliptic_workspace(elliptic_t* elliptic)
{
  int rc;
  if ((rc = elliptic->redPartials.alloc((size_t)kMaxRedBlocks * kMaxRed))) return rc;
  if ((rc = elliptic->redTicket.alloc(1))) return rc;
  elliptic->ws.partials = elliptic->redPartials.p;
  elliptic->ws.ticket = elliptic->redTicket.p;
  if (elliptic->comm && elliptic->comm->nranks > 1) elliptic->ws.peer = elliptic->comm->peerReduce();
  if ((rc = elliptic->o_scal.alloc(S_COUNT))) return rc;
  if (!elliptic->h_scal) NRSB_CUDA(cudaMallocHost((void**)&elliptic->h_scal, sizeof(double) * S_COUNT));
  if (!elliptic->h_err) {
    NRSB_CUDA(cudaHostAlloc((void**)&elliptic->h_err, sizeof(int), cudaHostAllocMapped));
    *elliptic->h_err = 0;
    NRSB_CUDA(cudaHostGetDevicePointer((void**)&elliptic->d_err, elliptic->h_err, 0));
  }
  return NRSB_OK;
}

// Krylov workspace that depends on SOLVER / PGMRES RESTART (ellipticSetup.cpp:225-262, GmresData in
// PGMRES.cpp:31-68).  Called from ellipticSolveSetup and again whenever nrsb_elliptic_set_option changes one of
// those keys (kershaw.udf:47-53 switches SOLVER between its benchmarks), so that pgmres() never runs on buffers
// sized for another configuration.
int ellipticKrylovWorkspace(elliptic_t* elliptic)
{
  options_t& options = elliptic->options;
  if (!options.compareArgs("SOLVER", "PGMRES")) return NRSB_OK;
  int rc;
  const size_t fo = (size_t)elliptic->fieldOffset * elliptic->Nfields;
  elliptic->nRestartVectors = 15;
  options.getArgs("PGMRES RESTART", elliptic->nRestartVectors);
  NRSB_REQUIRE(elliptic->nRestartVectors >= 1 && elliptic->nRestartVectors <= kMaxRed,
               "PGMRES RESTART must be in 1..16");
  const int m = elliptic->nRestartVectors;
  const bool flexible = options.compareArgs("SOLVER", "FLEXIBLE");
  NRSB_CUDA(cudaStreamSynchronize(elliptic->stream));  // buffers may be re-allocated: nothing may still use them
  if (elliptic->o_V.n < fo * m)
    if ((rc = elliptic->o_V.alloc(fo * m))) return rc;
  if (elliptic->o_Z.n < fo * (flexible ? m : 1))
    if ((rc = elliptic->o_Z.alloc(fo * (flexible ? m : 1)))) return rc;
  if (elliptic->o_y.n < (size_t)kMaxRed)
    if ((rc = elliptic->o_y.alloc(kMaxRed))) return rc;
  if (!flexible && elliptic->o_rtmp.n < fo)
    if ((rc = elliptic->o_rtmp.alloc(fo))) return rc;
  elliptic->gmres_H.assign((size_t)(m + 1) * (m + 1), 0.0);
  elliptic->gmres_sn.assign(m, 0.0);
  elliptic->gmres_cs.assign(m, 0.0);
  elliptic->gmres_s.assign(m + 1, 0.0);
  elliptic->gmres_y.assign(m, 0.0);
  NRSB_CUDA(cudaDeviceSynchronize());  // dbuf::alloc zero-fills on the legacy stream
  return NRSB_OK;
}

int ellipticSolveSetup(elliptic_t* elliptic)
{
  mesh_t* mesh = elliptic->mesh;
  options_t& options = elliptic->options;
  NRSB_REQUIRE(!elliptic->name.empty(), "Empty elliptic solver name!");
  options.setArgs("DISCRETIZATION", "CONTINUOUS");
  NRSB_REQUIRE(elliptic->Nfields >= 1 && elliptic->Nfields <= 3, "Invalid Nfields");  // ellipticSetup.cpp:81-86
  if (elliptic->Nfields > 1) {
    // block solver (ellipticSetup.cpp:53-57,131): Jacobi or no preconditioner, PCG, Helmholtz form, fp64
    NRSB_REQUIRE(!options.compareArgs("PRECONDITIONER", "MULTIGRID"),
                 "Block solver is implemented for C0-HEXES with Jacobi preconditioner only");
    NRSB_REQUIRE(options.compareArgs("SOLVER", "PCG"), "block solves use PCG");
    NRSB_REQUIRE(!options.compareArgs("INITIAL GUESS", "PROJECTION"), "block solves: no solution projection yet");
    NRSB_REQUIRE(elliptic->Nfields == 3, "block kernels are built for three fields");
    elliptic->poisson = false;
    if (elliptic->stressForm)
      if (int rcv = mesh->ensure_vgeo()) return rcv;
  }
  // fieldOffset: Nlocal rounded up to ALIGN_SIZE bytes (setup.cpp:295-299, nrssys.hpp:122)
  if (elliptic->fieldOffset == 0) {
    const dlong per = 1024 / sizeof(double);
    elliptic->fieldOffset = ((mesh->Nlocal + per - 1) / per) * per;
  }
  int rc;
  if ((rc = elliptic_workspace(elliptic))) return rc;
  const size_t fo = (size_t)elliptic->fieldOffset * elliptic->Nfields;
  if ((rc = elliptic->o_p.alloc(fo))) return rc;
  if ((rc = elliptic->o_z.alloc(fo))) return rc;
  if ((rc = elliptic->o_Ap.alloc(fo))) return rc;
  if ((rc = elliptic->o_x0.alloc(fo))) return rc;
  if ((rc = elliptic->o_rPfloat.alloc(fo))) return rc;
  if ((rc = elliptic->o_zPfloat.alloc(fo))) return rc;
  std::vector<double> l0(1, elliptic->lambda0Value), l1(1, elliptic->lambda1Value);
  if (elliptic->Nfields > 1) {
    // one constant per field, read by the block kernels at lambda[fld * loffset] with loffset = 1
    l0.assign(elliptic->Nfields, elliptic->lambda0Value);
    l1.assign(elliptic->Nfields, elliptic->lambda1Value);
    if (!elliptic->blockLambda0.empty()) l0 = elliptic->blockLambda0;
    if (!elliptic->blockLambda1.empty()) l1 = elliptic->blockLambda1;
    NRSB_REQUIRE((int)l0.size() == elliptic->Nfields && (int)l1.size() == elliptic->Nfields,
                 "block coefficients: one value per field");
    if (!elliptic->lambdaField) elliptic->loffset = 1;
  }
  if ((rc = elliptic->o_lambda0.upload(l0))) return rc;
  if ((rc = elliptic->o_lambda1.upload(l1))) return rc;
  std::vector<float> l0f(1, (float)elliptic->lambda0Value), l1f(1, (float)elliptic->lambda1Value);
  if ((rc = elliptic->o_lambda0Pfloat.upload(l0f))) return rc;
  if ((rc = elliptic->o_lambda1Pfloat.upload(l1f))) return rc;

  // allNeumann (ellipticSetup.cpp:182-204)
  elliptic->allNeumann = 0;
  if (elliptic->poisson) {
    int allNeumann = 1;
    for (int bc : elliptic->EToB)
      if (bc > 0 && bc != 4) allNeumann = 0;  // NEUMANN == 4
    if (elliptic->comm && elliptic->comm->nranks > 1) {
      std::vector<int> buf(elliptic->comm->nranks, 1);
      buf[elliptic->comm->rank] = allNeumann;
      elliptic->comm->allgather_bytes(buf.data(), sizeof(int));
      for (int v : buf) allNeumann = std::min(allNeumann, v);
    }
    elliptic->allNeumann = allNeumann;
  }

  if ((rc = ellipticOgs(mesh, elliptic->EToB, elliptic))) return rc;

  // ENABLE GS COMM OVERLAP: the reference times both variants and keeps the faster
  // (ellipticSetup.cpp:278-302).  Splitting only pays when there are halo rows.
  // Measured (2 B200, E = 4096 per GPU, N = 7 fp64): Ax on all 148 SMs + the one-launch flag-in-data exchange
  // 41.1 us per operator, the single launch with 8 pusher CTAs + finish 41.8 us (and the pushers take 15-16 SMs on
  // 4 and 8 GPUs): the in-kernel push is opt-in (FUSED HALO AX = TRUE) since the exchange became one launch.
  elliptic->fusedHaloAx = options.compareArgs("FUSED HALO AX", "TRUE") || getenv("NRSB_FUSED_HALO") != nullptr;
  // measured (tools/gs_timing.py, B200): phase 2 with 192 threads per SM needs 15 us for the rows the separate
  // 2048-threads-per-SM kernel does in 12.5 us (both bound by LSU wavefronts of the scattered 8-byte accesses),
  // 42.0 vs 39.4 us per operator at E=4096: off unless asked for
  elliptic->fusedGsAx = options.compareArgs("FUSED GS AX", "TRUE") || getenv("NRSB_FUSED_GS") != nullptr;
  elliptic->fusedDotAx = !options.compareArgs("FUSED DOT AX", "FALSE") && getenv("NRSB_NO_FUSED_DOT") == nullptr;
  {
    // the fused launches are persistent grids of kNumSMs co-resident CTAs that wait for each other: only on a device
    // that really has that many SMs to itself
    int dev = 0, sms = 0;
    NRSB_CUDA(cudaGetDevice(&dev));
    NRSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sms < kNumSMs) elliptic->fusedHaloAx = elliptic->fusedGsAx = false;
  }
  elliptic->overlap = elliptic->ogs->NhaloGather > 0 && !options.compareArgs("ENABLE GS COMM OVERLAP", "FALSE") &&
                      mesh->NlocalGatherElements > 0;

  if ((rc = ellipticKrylovWorkspace(elliptic))) return rc;
  if ((rc = ellipticChooseOverlap(elliptic, 8))) return rc;

  if ((rc = ellipticPreconditionerSetup(elliptic))) return rc;

  if (options.compareArgs("INITIAL GUESS", "PROJECTION")) {
    int nVecsProject = 8, nStepsStart = 5;
    options.getArgs("RESIDUAL PROJECTION VECTORS", nVecsProject);
    options.getArgs("RESIDUAL PROJECTION START", nStepsStart);
    const bool aconj = options.compareArgs("INITIAL GUESS", "PROJECTION-ACONJ");
    elliptic->solutionProjection.reset(new SolutionProjection(elliptic, aconj, nVecsProject, nStepsStart));
    if ((rc = elliptic->solutionProjection->setup())) return rc;
  }
  NRSB_CUDA(cudaDeviceSynchronize());
  return NRSB_OK;
}

int ellipticPreconditionerSetup(elliptic_t* elliptic)
{
  options_t& options = elliptic->options;
  elliptic->precon.reset(new precon_t());
  if (options.compareArgs("PRECONDITIONER", "MULTIGRID")) return ellipticMultiGridSetup(elliptic, elliptic->precon.get());
  if (options.compareArgs("PRECONDITIONER", "JACOBI")) {
    int rc;
    if ((rc = elliptic->precon->o_invDiagA.alloc((size_t)elliptic->fieldOffset * elliptic->Nfields))) return rc;
    return ellipticBuildDiagonal<double>(elliptic, elliptic->precon->o_invDiagA.p);
  }
  if (options.compareArgs("PRECONDITIONER", "NONE") || !options.has("PRECONDITIONER")) {
    options.setArgs("PRECONDITIONER", "NONE");
    return NRSB_OK;
  }
  set_last_error("unsupported PRECONDITIONER '" + options.getArgs("PRECONDITIONER") +
                 "' (supported: MULTIGRID, JACOBI, NONE)");
  return NRSB_ERR_INVALID;
}

int ellipticPreconditioner(elliptic_t* elliptic, double* o_r, double* o_z)
{
  mesh_t* mesh = elliptic->mesh;
  options_t& options = elliptic->options;
  precon_t* precon = elliptic->precon.get();
  const long Nall = (long)elliptic->fieldOffset * elliptic->Nfields;
  int rc;
  if (options.compareArgs("PRECONDITIONER", "JACOBI")) {
    if ((rc = axmyz_launch<double>(elliptic->Nvec(), 1.0, o_r, precon->o_invDiagA.p, o_z, elliptic->stream))) return rc;
  } else if (options.compareArgs("PRECONDITIONER", "MULTIGRID")) {
    // pfill(z)=0 ; cast r ; V-cycle ; cast z  (ellipticPreconditioner.cpp:57-62)
    if ((rc = fill_launch<float>(Nall, 0.f, elliptic->o_zPfloat.p, elliptic->stream))) return rc;
    if ((rc = copy_d2f_launch(Nall, o_r, elliptic->o_rPfloat.p, elliptic->stream))) return rc;
    if ((rc = precon->MGSolver->Run(elliptic->o_rPfloat.p, elliptic->o_zPfloat.p))) return rc;
    if ((rc = copy_f2d_launch(Nall, elliptic->o_zPfloat.p, o_z, elliptic->stream))) return rc;
  } else if (options.compareArgs("PRECONDITIONER", "NONE")) {
    NRSB_CUDA(cudaMemcpyAsync(o_z, o_r, sizeof(double) * Nall, cudaMemcpyDeviceToDevice, elliptic->stream));
  } else {
    set_last_error("Unknown preconditioner");
    return NRSB_ERR_INVALID;
  }
  if (elliptic->allNeumann) return ellipticZeroMean(elliptic, o_z);
  return NRSB_OK;
}

// ------------------------------------------------------------------------------------------
// PCG (PCG.cpp:85-203)
// ------------------------------------------------------------------------------------------
int pcg(elliptic_t* elliptic, double* o_r, double* o_x, double tol, int MAXIT, double& rdotr)
{
  mesh_t* mesh = elliptic->mesh;
  options_t& options = elliptic->options;
  const bool flexible = options.compareArgs("SOLVER", "FLEXIBLE");
  const bool precond = !options.compareArgs("PRECONDITIONER", "NONE");
  cudaStream_t st = elliptic->stream;
  double* S = elliptic->o_scal.p;
  double* o_p = elliptic->o_p.p;
  double* o_z = precond ? elliptic->o_z.p : o_r;
  double* o_Ap = elliptic->o_Ap.p;
  const double* o_weight = elliptic->o_weight();
  const long N = elliptic->Nvec();
  int rc;
  if ((rc = fill_launch<double>((long)elliptic->fieldOffset * elliptic->Nfields, 0.0, o_p, st))) return rc;

  // rdotr enters as res0Norm = sqrt(sum w r^2 * resNormFactor); without preconditioner the first
  // rdotz1 is `rdotr` itself (PCG.cpp:131: "rdotz1 = rdotr"), i.e. the NORM, not its square -- a
  // quirk of the reference that cancels in alpha*p for iteration 1 only through beta = 0; it is
  // reproduced literally.
  double h_rdotr = rdotr;
  NRSB_CUDA(cudaMemcpyAsync(S + S_RDOTR, &h_rdotr, sizeof(double), cudaMemcpyHostToDevice, st));
  NRSB_CUDA(cudaStreamSynchronize(st));

  // The loop runs ahead of the host: the residual norm, its history and the convergence flag live on the
  // device (PcgControl); the host looks at the flag every `interval` iterations only.  Once the flag is up,
  // x and r are frozen, so the solution and the iteration count are exactly those of the reference's
  // check-every-iteration loop (PCG.cpp:115-200); the iterations launched in between are wasted work,
  // which is why the interval is 1 when an iteration is expensive (multigrid).
  int interval = options.compareArgs("PRECONDITIONER", "MULTIGRID") ? 1 : 8;
  options.getArgs("PCG CHECK INTERVAL", interval);
  interval = std::max(1, interval);
  if (elliptic->o_resHist.n < (size_t)MAXIT + 1)
    if ((rc = elliptic->o_resHist.alloc((size_t)MAXIT + 1))) return rc;
  if (!elliptic->o_dotPartials.p)
    if ((rc = elliptic->o_dotPartials.alloc(kNumSMs))) return rc;
  {
    const double init[3] = {0.0, 0.0, rdotr};
    NRSB_CUDA(cudaMemcpyAsync(S + S_DONE, init, sizeof(init), cudaMemcpyHostToDevice, st));
    NRSB_CUDA(cudaStreamSynchronize(st));
  }
  PcgControl ctl;
  ctl.ctl = S + S_DONE;
  ctl.hist = elliptic->o_resHist.p;
  ctl.factor = elliptic->resNormFactor;
  ctl.tol = tol;
  ctl.rdotz = S + S_RDOTZ;
  ctl.rdotzOld = S + S_RDOTZ_OLD;
  ctl.normIsRdotz = precond ? 0 : 1;
  if (!precond)  // first iteration: rdotz1 = rdotr (the norm on entry)
    NRSB_CUDA(cudaMemcpyAsync(S + S_RDOTZ, S + S_RDOTR, sizeof(double), cudaMemcpyDeviceToDevice, st));

  int iter = 0;
  bool done = false;
  do {
    iter++;
    // rdotz2 = rdotz1 was done by the last block of the previous iteration's update kernel (PcgControl), and so
    // was, without preconditioner, rdotz1 = rdotr: the residual NORM of the previous iteration
    if (precond) {
      if ((rc = ellipticPreconditioner(elliptic, o_r, o_z))) return rc;
      if ((rc = wdot_launch<double>(N, o_weight, o_r, o_z, S + S_RDOTZ, elliptic->ws, st))) return rc;
    }
    DevScalar beta = DevScalar::host(0.0);
    if (iter > 1) {
      beta = DevScalar::ratio(S + S_RDOTZ, S + S_RDOTZ_OLD);
      if (flexible) {
        if ((rc = wdot_launch<double>(N, o_weight, o_z, o_Ap, S + S_ZDOTAP, elliptic->ws, st))) return rc;
        // beta = -alpha * zdotAp / rdotz2   (alpha of the previous iteration is still in S_ALPHA)
        DevScalar b = DevScalar::ratio(S + S_ZDOTAP, S + S_RDOTZ_OLD, -1.0);
        b.mul = S + S_ALPHA;
        if ((rc = set_scalar_launch(S + S_BETA, b, st))) return rc;
        beta = DevScalar::ratio(S + S_BETA, nullptr);
      }
    }
    // p = z + beta p
    if ((rc = axpby_launch<double>(N, DevScalar::host(1.0), o_z, beta, o_p, st))) return rc;
    AxDot dot;
    dot.partials = elliptic->o_dotPartials.p;
    if ((rc = ellipticOperator<double>(elliptic, o_p, o_Ap, true, elliptic->fusedDotAx ? &dot : nullptr))) return rc;
    // pAp and alpha = rdotz1 / (pAp + 1e-300) in one launch: either the fold of the per-CTA values of p^T A_L p
    // that the axhelm launch left behind (p is continuous and masked, so this IS sum invDegree p Ap), or the
    // reference's separate pass over p and Ap (PCG.cpp:150-157)
    if (dot.n > 0)
      rc = sum_ratio_launch(dot.n, dot.partials, S + S_PAP, S + S_RDOTZ, 1e-300, S + S_ALPHA, elliptic->ws, st);
    else
      rc = wdot_ratio_launch(N, o_weight, o_p, o_Ap, S + S_PAP, S + S_RDOTZ, 1e-300, S + S_ALPHA, elliptic->ws, st);
    if (rc) return rc;
    DevScalar alpha = DevScalar::ratio(S + S_ALPHA, nullptr);
    // x += alpha p ; r -= alpha Ap ; rdotr = sqrt(sum w r^2 * factor) ; history ; convergence flag  (one kernel)
    ctl.iter = iter;
    if ((rc = update_pcg_ctl_launch(N, o_weight, o_Ap, o_p, alpha, o_r, o_x, S + S_NORM, ctl, elliptic->ws, st)))
      return rc;
    if (iter % interval == 0 || iter == MAXIT) {
      double v[2];
      if ((rc = elliptic->read_scalars(S_DONE, 2, v))) return rc;
      if (v[0] != 0.0) {
        done = true;
        iter = (int)v[1];
      }
    }
  } while (!done && iter < MAXIT);
  {
    std::vector<double> h(iter);
    NRSB_CUDA(cudaMemcpyAsync(h.data(), elliptic->o_resHist.p, sizeof(double) * iter, cudaMemcpyDeviceToHost, st));
    NRSB_CUDA(cudaStreamSynchronize(st));
    for (double v : h) elliptic->resHistory.push_back(v);
    if (iter > 0) rdotr = h[iter - 1];
  }
  if (std::isnan(rdotr)) {
    set_last_error("Detected invalid resiual norm while running linear solver!");
    return NRSB_ERR_DIVERGED;
  }
  elliptic->Niter = iter;
  return NRSB_OK;
}

// ------------------------------------------------------------------------------------------
// PGMRES (PGMRES.cpp:70-340)
// ------------------------------------------------------------------------------------------
static int gmresUpdate(elliptic_t* elliptic, double* o_x, int gmresUpdateSize)
{
  const int m = elliptic->nRestartVectors;
  mesh_t* mesh = elliptic->mesh;
  std::vector<double>&y = elliptic->gmres_y, &H = elliptic->gmres_H, &s = elliptic->gmres_s;
  cudaStream_t st = elliptic->stream;
  const long fo = (long)elliptic->fieldOffset * elliptic->Nfields;
  for (int k = gmresUpdateSize - 1; k >= 0; --k) {
    y[k] = s[k];
    for (int j = k + 1; j < gmresUpdateSize; ++j) y[k] -= H[k + j * (m + 1)] * y[j];
    y[k] /= H[k + k * (m + 1)];
  }
  NRSB_CUDA(cudaMemcpyAsync(elliptic->o_y.p, y.data(), gmresUpdateSize * sizeof(double), cudaMemcpyHostToDevice, st));
  NRSB_CUDA(cudaStreamSynchronize(st));
  int rc;
  if (elliptic->options.compareArgs("SOLVER", "FLEXIBLE"))
    return update_pgmres_solution_launch(mesh->Nlocal, fo, gmresUpdateSize, elliptic->o_y.p, elliptic->o_Z.p, o_x, st);
  if ((rc = fill_launch<double>(fo, 0.0, elliptic->o_z.p, st))) return rc;
  if ((rc = update_pgmres_solution_launch(mesh->Nlocal, fo, gmresUpdateSize, elliptic->o_y.p, elliptic->o_V.p,
                                          elliptic->o_z.p, st)))
    return rc;
  if ((rc = ellipticPreconditioner(elliptic, elliptic->o_z.p, elliptic->o_p.p))) return rc;
  return axpby_launch<double>(mesh->Nlocal, DevScalar::host(1.0), elliptic->o_p.p, DevScalar::host(1.0), o_x, st);
}

int pgmres(elliptic_t* elliptic, double* o_r, double* o_x, double tol, int MAXIT, double& rdotr)
{
  mesh_t* mesh = elliptic->mesh;
  cudaStream_t st = elliptic->stream;
  const int m = elliptic->nRestartVectors;
  const bool flexible = elliptic->options.compareArgs("SOLVER", "FLEXIBLE");
  const long fo = (long)elliptic->fieldOffset * elliptic->Nfields;
  const long N = mesh->Nlocal;
  NRSB_REQUIRE(m >= 1 && elliptic->o_V.n >= (size_t)fo * m && elliptic->o_Z.n >= (size_t)fo * (flexible ? m : 1) &&
                   elliptic->o_y.p && (flexible || elliptic->o_rtmp.n >= (size_t)fo) &&
                   elliptic->gmres_H.size() == (size_t)(m + 1) * (m + 1),
               "pgmres: workspace was not set up for this SOLVER / PGMRES RESTART (ellipticKrylovWorkspace)");
  double* o_w = elliptic->o_p.p;
  double* o_Ax = elliptic->o_Ap.p;
  double* o_V = elliptic->o_V.p;
  double* o_Z = elliptic->o_Z.p;
  // The reference aliases b with elliptic->o_z (PGMRES.cpp:128), which the NON-flexible gmresUpdate
  // overwrites (PGMRES.cpp:83-96), so its restarted non-flexible GMRES computes r = b - Ax from a
  // clobbered b.  b is kept in its own buffer here when not flexible (deviation noted in DESIGN.md).
  double* o_b = flexible ? elliptic->o_z.p : elliptic->o_rtmp.p;
  double* S = elliptic->o_scal.p;
  const double* o_weight = elliptic->o_invDegree;
  std::vector<double>&y = elliptic->gmres_y, &H = elliptic->gmres_H, &sn = elliptic->gmres_sn, &cs = elliptic->gmres_cs,
  &s = elliptic->gmres_s;
  int rc;
  NRSB_CUDA(cudaMemcpyAsync(o_b, o_r, sizeof(double) * fo, cudaMemcpyDeviceToDevice, st));
  double nr = rdotr / std::sqrt(elliptic->resNormFactor);
  double error = rdotr;
  const double TOL = tol;
  int iter = 0;
  for (iter = 0; iter < MAXIT;) {
    s[0] = nr;
    // V(:,0) = r/nr
    if ((rc = axpby_launch<double>(N, DevScalar::host(1.0 / nr), o_r, DevScalar::host(0.0), o_V, st))) return rc;
    for (int i = 0; i < m; ++i) {
      double* o_Mv = flexible ? o_Z + (size_t)i * fo : o_Z;
      if ((rc = ellipticPreconditioner(elliptic, o_V + (size_t)i * fo, o_Mv))) return rc;
      if ((rc = ellipticOperator<double>(elliptic, o_Mv, o_w))) return rc;
      // y = V^T w (weighted): stays on the device for the Gram-Schmidt kernel, one copy to the host for H
      if ((rc = wdot_multi_launch(N, i + 1, fo, o_weight, o_V, o_w, elliptic->o_y.p, elliptic->ws, st))) return rc;
      if ((rc = gram_schmidt_launch(N, fo, i + 1, o_weight, elliptic->o_y.p, o_V, o_w, S + S_NORM, elliptic->ws, st)))
        return rc;
      NRSB_CUDA(cudaMemcpyAsync(S + S_GMRES, elliptic->o_y.p, sizeof(double) * (i + 1), cudaMemcpyDeviceToDevice, st));
      NRSB_CUDA(cudaMemcpyAsync(S + S_GMRES + (i + 1), S + S_NORM, sizeof(double), cudaMemcpyDeviceToDevice, st));
      std::vector<double> hv(i + 2);
      if ((rc = elliptic->read_scalars(S_GMRES, i + 2, hv.data()))) return rc;
      for (int k = 0; k <= i; ++k) y[k] = hv[k];
      const double nw = std::sqrt(hv[i + 1]);
      H[i + 1 + i * (m + 1)] = nw;
      if (i < m - 1)
        if ((rc = axpby_launch<double>(N, DevScalar::host(1. / nw), o_w, DevScalar::host(0.0),
                                       o_V + (size_t)(i + 1) * fo, st)))
          return rc;
      for (int k = 0; k <= i; ++k) H[k + i * (m + 1)] = y[k];
      for (int k = 0; k < i; ++k) {
        const double h1 = H[k + i * (m + 1)], h2 = H[k + 1 + i * (m + 1)];
        H[k + i * (m + 1)] = cs[k] * h1 + sn[k] * h2;
        H[k + 1 + i * (m + 1)] = -sn[k] * h1 + cs[k] * h2;
      }
      const double h1 = H[i + i * (m + 1)], h2 = H[i + 1 + i * (m + 1)];
      const double hr = std::sqrt(h1 * h1 + h2 * h2);
      cs[i] = h1 / hr;
      sn[i] = h2 / hr;
      H[i + i * (m + 1)] = cs[i] * h1 + sn[i] * h2;
      H[i + 1 + i * (m + 1)] = 0;
      s[i + 1] = -sn[i] * s[i];
      s[i] = cs[i] * s[i];
      iter++;
      error = std::fabs(s[i + 1]) * std::sqrt(elliptic->resNormFactor);
      rdotr = error;
      elliptic->resHistory.push_back(rdotr);
      if (std::isnan(error)) {
        set_last_error("Detected invalid resiual norm while running linear solver!");
        return NRSB_ERR_DIVERGED;
      }
      if (error < TOL || iter == MAXIT) {
        if ((rc = gmresUpdate(elliptic, o_x, i + 1))) return rc;
        break;
      }
    }
    if (error < TOL || iter == MAXIT) break;
    if ((rc = gmresUpdate(elliptic, o_x, m))) return rc;
    if ((rc = ellipticOperator<double>(elliptic, o_x, o_Ax))) return rc;
    if ((rc = fused_residual_and_norm_launch(N, o_weight, o_b, o_Ax, o_r, S + S_NORM, elliptic->ws, st))) return rc;
    double v;
    if ((rc = elliptic->read_scalars(S_NORM, 1, &v))) return rc;
    nr = std::sqrt(v);
    error = nr * std::sqrt(elliptic->resNormFactor);
    rdotr = error;
    if (error <= TOL) break;
  }
  elliptic->Niter = iter;
  return NRSB_OK;
}

// ------------------------------------------------------------------------------------------
// ellipticSolve (ellipticSolve.cpp:32-190)
// ------------------------------------------------------------------------------------------
static int weighted_norm(elliptic_t* elliptic, const double* o_v, double* out)
{
  int rc = wnorm2_launch<double>(elliptic->Nvec(), elliptic->o_weight(), o_v, elliptic->o_scal.p + S_NORM,
                                 elliptic->ws, elliptic->stream);
  if (rc) return rc;
  double v;
  if ((rc = elliptic->read_scalars(S_NORM, 1, &v))) return rc;
  *out = std::sqrt(v) * std::sqrt(elliptic->resNormFactor);
  return NRSB_OK;
}

int ellipticSolve(elliptic_t* elliptic, double* o_r, double* o_x)
{
  options_t& options = elliptic->options;
  mesh_t* mesh = elliptic->mesh;
  cudaStream_t st = elliptic->stream;
  const long N = elliptic->Nvec();
  const long fo = (long)elliptic->fieldOffset * elliptic->Nfields;
  int maxIter = 999;
  options.getArgs("MAXIMUM ITERATIONS", maxIter);
  elliptic->resNormFactor = 1 / mesh->volume;
  elliptic->resHistory.clear();
  int rc;

  // coefficient fields that change between solves: refresh the preconditioner's copies (ellipticSolve.cpp:79-88)
  if (options.compareArgs("ELLIPTIC PRECO COEFF FIELD", "TRUE")) {
    if (options.compareArgs("PRECONDITIONER", "MULTIGRID"))
      if ((rc = ellipticMultiGridUpdateLambda(elliptic))) return rc;
    if (options.compareArgs("PRECONDITIONER", "JACOBI") || options.compareArgs("MULTIGRID SMOOTHER", "DAMPEDJACOBI"))
      if ((rc = ellipticUpdateJacobi(elliptic))) return rc;
  }

  // r = rhs - A x0
  if ((rc = ellipticAx<double>(elliptic, mesh->Nelements, mesh->o_elementList.p, o_x, elliptic->o_Ap.p))) return rc;
  if ((rc = axpby_launch<double>(N, DevScalar::host(-1.0), elliptic->o_Ap.p, DevScalar::host(1.0), o_r, st))) return rc;
  if (elliptic->allNeumann)
    if ((rc = ellipticZeroMean(elliptic, o_r))) return rc;
  // mask + gather-scatter of the residual
  if (elliptic->oogs->ogs->NhaloGather || elliptic->Nfields > 1) {  // (block solver: unmasked numbering)
    if ((rc = ellipticApplyMask<double>(elliptic, o_r))) return rc;
    if ((rc = elliptic->oogs->startFinish<double>(o_r, elliptic->Nfields, elliptic->fieldOffset, gs_op::add, 0, nullptr,
                                                  st)))
      return rc;
  } else if ((rc = elliptic->oogs->startFinish<double>(o_r, elliptic->Nfields, elliptic->fieldOffset, gs_op::add,
                                                       elliptic->Nmasked, elliptic->o_maskIds.p, st)))
    return rc;

  NRSB_CUDA(cudaMemcpyAsync(elliptic->o_x0.p, o_x, sizeof(double) * fo, cudaMemcpyDeviceToDevice, st));
  if ((rc = fill_launch<double>(fo, 0.0, o_x, st))) return rc;
  const bool projection = options.compareArgs("INITIAL GUESS", "PROJECTION") && elliptic->solutionProjection;
  if (projection) {
    if ((rc = weighted_norm(elliptic, o_r, &elliptic->res00Norm))) return rc;
    if (std::isnan(elliptic->res00Norm)) {
      set_last_error(elliptic->name + " unreasonable res00Norm!");
      return NRSB_ERR_DIVERGED;
    }
    if ((rc = elliptic->solutionProjection->pre(o_r))) return rc;
  }
  if ((rc = weighted_norm(elliptic, o_r, &elliptic->res0Norm))) return rc;
  if (std::isnan(elliptic->res0Norm)) {
    set_last_error(elliptic->name + " unreasonable res00Norm!");
    return NRSB_ERR_DIVERGED;
  }
  double tol = 1e-6;
  options.getArgs("SOLVER TOLERANCE", tol);
  if (options.compareArgs("LINEAR SOLVER STOPPING CRITERION", "RELATIVE")) tol *= elliptic->res0Norm;

  elliptic->resNorm = elliptic->res0Norm;
  if (options.compareArgs("SOLVER", "PCG")) {
    if ((rc = pcg(elliptic, o_r, o_x, tol, maxIter, elliptic->resNorm))) return rc;
  } else if (options.compareArgs("SOLVER", "PGMRES")) {
    if ((rc = pgmres(elliptic, o_r, o_x, tol, maxIter, elliptic->resNorm))) return rc;
  } else {
    set_last_error("Linear solver " + options.getArgs("SOLVER") + " is not supported!");
    return NRSB_ERR_INVALID;
  }
  if (projection) {
    if ((rc = elliptic->solutionProjection->post(o_x))) return rc;
  } else {
    elliptic->res00Norm = elliptic->res0Norm;
  }
  // x += x0
  if ((rc = axpby_launch<double>(N, DevScalar::host(1.0), elliptic->o_x0.p, DevScalar::host(1.0), o_x, st))) return rc;
  if (elliptic->allNeumann)
    if ((rc = ellipticZeroMean(elliptic, o_x))) return rc;
  NRSB_CUDA(cudaStreamSynchronize(st));
  return NRSB_OK;
}

}  // namespace nrsb

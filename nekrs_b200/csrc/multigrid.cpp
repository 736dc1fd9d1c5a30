// multigrid.cpp -- p-multigrid preconditioner: level construction, V-cycle, Chebyshev smoothers,
// overlapping Schwarz (FDM) setup and application, coarse solve.
//
// Restates: determineMGLevels.cpp:5-98, ellipticMultiGridSetup.cpp:60-343,
// ellipticBuildMultigridLevel(Fine).cpp, createMeshMG (meshSetup.cpp:293-376), MGSolver.cpp:150-193,
// ellipticMultiGridLevel.cpp:32-277, ellipticMultiGridLevelSetup.cpp:109-180,238-258,292-453,
// ellipticMultiGridSchwarz.cpp:58-1156, coarseLevel.cpp:182-222.
//
// Not restated (documented in DESIGN.md): the BoomerAMG coarse solve (third-party hypre, CPU).  Its
// place is taken by a device-resident Jacobi-PCG on the same assembled N=1 operator
// (`COARSE SOLVER = JPCG`); `COARSE SOLVER = SMOOTHER` (reference-supported, parReader.cpp:778-781)
// is implemented literally.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "host.hpp"
#include "optimal_coeffs.hpp"

namespace nrsb {

int elliptic_workspace(elliptic_t* elliptic);

// ------------------------------------------------------------------------------------------
std::vector<int> determineMGLevels(const options_t& options, int N)
{
  // user schedule "p=7+degree=3, p=3+degree=3, p=1, ..." (parseMultigridSchedule): only the orders matter here
  // (option values are upper-cased on the way in, capi_elliptic.cu parse_options)
  std::string schedule = options.getArgs("MULTIGRID SCHEDULE");
  for (auto& c : schedule) c = (char)tolower(c);
  if (!schedule.empty()) {
    std::vector<int> levels;
    size_t pos = 0;
    while ((pos = schedule.find("p=", pos)) != std::string::npos) {
      const int p = std::atoi(schedule.c_str() + pos + 2);
      if (std::find(levels.begin(), levels.end(), p) == levels.end()) levels.push_back(p);
      pos += 2;
    }
    std::sort(levels.rbegin(), levels.rend());
    return levels;
  }
  static const std::map<int, std::vector<int>> schwarz = {
      {1, {1}},        {2, {2, 1}},     {3, {3, 1}},     {4, {4, 2, 1}},  {5, {5, 3, 1}},
      {6, {6, 3, 1}},  {7, {7, 3, 1}},  {8, {8, 5, 1}},  {9, {9, 5, 1}},  {10, {10, 6, 1}},
      {11, {11, 6, 1}}};
  static const std::map<int, std::vector<int>> other = {
      {1, {1}},           {2, {2, 1}},        {3, {3, 1}},        {4, {4, 2, 1}},      {5, {5, 3, 1}},
      {6, {6, 4, 2, 1}},  {7, {7, 5, 3, 1}},  {8, {8, 6, 4, 1}},  {9, {9, 7, 5, 1}},   {10, {10, 8, 5, 1}},
      {11, {11, 9, 5, 1}}};
  if (options.compareArgs("MULTIGRID SMOOTHER", "ASM") || options.compareArgs("MULTIGRID SMOOTHER", "RAS"))
    return schwarz.at(N);
  return other.at(N);
}

// degree of a (order, leg) pair in a user schedule; -1 if absent.  Schedule entries are listed fine
// to coarse (down leg) and back (up leg); "p=1" alone has no smoothing degree.
static int schedule_degree(const std::string& schedule_, int order, bool downLeg)
{
  std::string schedule = schedule_;
  for (auto& c : schedule) c = (char)tolower(c);
  std::vector<std::pair<int, int>> entries;
  size_t pos = 0;
  while ((pos = schedule.find("p=", pos)) != std::string::npos) {
    const int p = std::atoi(schedule.c_str() + pos + 2);
    size_t end = schedule.find(',', pos);
    if (end == std::string::npos) end = schedule.size();
    const std::string item = schedule.substr(pos, end - pos);
    int deg = -1;
    const size_t d = item.find("degree=");
    if (d != std::string::npos) deg = std::atoi(item.c_str() + d + 7);
    entries.push_back({p, deg});
    pos = end;
  }
  int minP = 1 << 30, minIdx = -1;
  for (size_t i = 0; i < entries.size(); ++i)
    if (entries[i].first < minP) {
      minP = entries[i].first;
      minIdx = (int)i;
    }
  for (size_t i = 0; i < entries.size(); ++i) {
    const bool isDown = (int)i <= minIdx;
    if (entries[i].first == order && isDown == downLeg) return entries[i].second;
  }
  return -1;
}

// ------------------------------------------------------------------------------------------
// level operators (ellipticMultiGridLevel.cpp)
// ------------------------------------------------------------------------------------------
int pMGLevel::Ax(const float* x, float* Ax_) { return ellipticOperator<float>(elliptic, x, Ax_); }

int pMGLevel::residual(const float* rhs, const float* x, float* res)
{
  int rc = ellipticOperator<float>(elliptic, x, res);
  if (rc) return rc;
  // res = rhs - res
  return axpby_launch<float>(Nrows, DevScalar::host(1.0), rhs, DevScalar::host(-1.0), res, elliptic->stream);
}

int pMGLevel::coarsen(float* x, float* Rx)
{
  cudaStream_t st = elliptic->stream;
  int rc;
  // x *= invDegreeFine  (paxmy, :46)
  if ((rc = axmyz_launch<float>((long)mesh->Nelements * NpF, 1.0f, o_invDegreeFine, x, x, st))) return rc;
  if ((rc = transfer_dispatch(true, NqF, mesh->Nq, mesh->Nelements, R.data(), x, Rx, st))) return rc;
  // gather-scatter + mask again (coarsen does not preserve the mask; the mask rides along in the same launch)
  return elliptic->oogs->startFinish<float>(Rx, 1, 0, gs_op::add, elliptic->Nmasked, elliptic->o_maskIds.p, st);
}

int pMGLevel::prolongate(const float* x, float* Px)
{
  return transfer_dispatch(false, NqF, mesh->Nq, mesh->Nelements, R.data(), x, Px, elliptic->stream);
}

int pMGLevel::smooth(const float* rhs, float* x, bool xIsZero)
{
  if (!xIsZero && (smootherType == SmootherType::ASM || smootherType == SmootherType::RAS)) return NRSB_OK;
  switch (smootherType) {
    case SmootherType::CHEBYSHEV: return smoothChebyshev(rhs, x, xIsZero);
    case SmootherType::OPT_FOURTH_CHEBYSHEV:
    case SmootherType::FOURTH_CHEBYSHEV: return smoothFourthKindChebyshev(rhs, x, xIsZero);
    case SmootherType::ASM:
    case SmootherType::RAS: return smoothSchwarz(rhs, x, xIsZero);
    case SmootherType::JACOBI: return smoothJacobi(rhs, x, xIsZero);
  }
  return NRSB_OK;
}

int pMGLevel::smoother(const float* x, float* Sx, bool xIsZero)
{
  if (chebySmootherType == ChebyshevSmootherType::JACOBI)
    return axmyz_launch<float>(Nrows, 1.0f, o_invDiagA.p, x, Sx, elliptic->stream);
  return smoothSchwarz(x, Sx, true);
}

int pMGLevel::smoothJacobi(const float* r, float* x, bool xIsZero)
{
  cudaStream_t st = elliptic->stream;
  float *res = o_smootherResidual.p, *d = o_smootherUpdate.p;
  int rc;
  if (xIsZero) return axmyz_launch<float>(Nrows, 1.0f, o_invDiagA.p, r, x, st);
  if ((rc = Ax(x, res))) return rc;
  if ((rc = axpby_launch<float>(Nrows, DevScalar::host(1.0), r, DevScalar::host(-1.0), res, st))) return rc;
  if ((rc = axmyz_launch<float>(Nrows, 1.0f, o_invDiagA.p, res, d, st))) return rc;
  return axpby_launch<float>(Nrows, DevScalar::host(1.0), d, DevScalar::host(1.0), x, st);
}

int pMGLevel::smoothChebyshev(const float* r, float* x, bool xIsZero)
{
  const int ChebyshevDegree = xIsZero ? DownLegChebyshevDegree : UpLegChebyshevDegree;
  if (ChebyshevDegree == 0) return NRSB_OK;
  cudaStream_t st = elliptic->stream;
  const float theta = 0.5 * (lambda1 + lambda0);
  const float delta = 0.5 * (lambda1 - lambda0);
  const float invTheta = 1.0 / theta;
  const float sigma = theta / delta;
  float rho_n = 1. / sigma;
  float *res = o_smootherResidual.p, *Ad = o_smootherResidual2.p, *d = o_smootherUpdate.p;
  int rc;
  if (xIsZero)
    if ((rc = fill_launch<float>(Nrows, 0.f, x, st))) return rc;
  if (!xIsZero) {
    if ((rc = Ax(x, res))) return rc;
    if ((rc = axpby_launch<float>(Nrows, DevScalar::host(1.0), r, DevScalar::host(-1.0), res, st))) return rc;
  } else {
    NRSB_CUDA(cudaMemcpyAsync(res, r, sizeof(float) * Nrows, cudaMemcpyDeviceToDevice, st));
  }
  if ((rc = smoother(res, res, xIsZero))) return rc;
  // d = invTheta * res
  if ((rc = axpby_launch<float>(Nrows, DevScalar::host(invTheta), res, DevScalar::host(0.0), d, st))) return rc;
  for (int k = 1; k < ChebyshevDegree; k++) {
    if ((rc = Ax(d, Ad))) return rc;
    if ((rc = smoother(Ad, Ad, xIsZero))) return rc;
    const float rhoSave = rho_n;
    rho_n = 1.0 / (2.0 * sigma - rho_n);
    const float rCoeff = 2.0 * rho_n / delta;
    const float dCoeff = rho_n * rhoSave;
    if ((rc = update_chebyshev_launch(Nrows, dCoeff, rCoeff, Ad, d, res, x, st))) return rc;
  }
  if ((rc = axpby_launch<float>(Nrows, DevScalar::host(1.0), d, DevScalar::host(1.0), x, st))) return rc;
  return ellipticApplyMask<float>(elliptic, x);
}

int pMGLevel::smoothFourthKindChebyshev(const float* r, float* x, bool xIsZero)
{
  const int ChebyshevDegree = xIsZero ? DownLegChebyshevDegree : UpLegChebyshevDegree;
  const std::vector<float>& betas = xIsZero ? DownLegBetas : UpLegBetas;
  if (ChebyshevDegree == 0) return NRSB_OK;
  cudaStream_t st = elliptic->stream;
  float *res = o_smootherResidual.p, *Ad = o_smootherResidual2.p, *d = o_smootherUpdate.p;
  const float rho = (float)this->lambda1;
  int rc;
  if (xIsZero) {
    if ((rc = fill_launch<float>(Nrows, 0.f, x, st))) return rc;
    NRSB_CUDA(cudaMemcpyAsync(res, r, sizeof(float) * Nrows, cudaMemcpyDeviceToDevice, st));
  } else {
    if ((rc = Ax(x, res))) return rc;
    if ((rc = axpby_launch<float>(Nrows, DevScalar::host(1.0), r, DevScalar::host(-1.0), res, st))) return rc;
  }
  // d = 4/(3 rho) S r
  if ((rc = smoother(res, Ad, xIsZero))) return rc;
  const float coeff = 4.0 / (3.0 * rho);
  if ((rc = axpby_launch<float>(Nrows, DevScalar::host(coeff), Ad, DevScalar::host(0.0), d, st))) return rc;
  for (int k = 1; k < ChebyshevDegree; k++) {
    if ((rc = Ax(d, Ad))) return rc;
    // x += beta d ; r -= Ad
    if ((rc = update_fourth_chebyshev_launch(Nrows, betas[k - 1], Ad, d, res, x, st))) return rc;
    if ((rc = smoother(res, Ad, xIsZero))) return rc;
    const float dCoeff = (2.0 * k - 1.0) / (2.0 * k + 3.0);
    const float rCoeff = (8.0 * k + 4.0) / ((2.0 * k + 3.0) * rho);
    if ((rc = axpby_launch<float>(Nrows, DevScalar::host(rCoeff), Ad, DevScalar::host(dCoeff), d, st))) return rc;
  }
  if ((rc = axpby_launch<float>(Nrows, DevScalar::host(betas.back()), d, DevScalar::host(1.0), x, st))) return rc;
  return ellipticApplyMask<float>(elliptic, x);
}

// smoothSchwarz (ellipticMultiGridSchwarz.cpp:1056-1156)
int pMGLevel::smoothSchwarz(const float* u, float* Su, bool /*xIsZero*/)
{
  cudaStream_t st = elliptic->stream;
  const dlong E = mesh->Nelements;
  int rc;
  if ((rc = pre_fdm_launch(mesh->Nq, E, u, o_work1.p, st))) return rc;
  if ((rc = ogsExt->startFinish<float>(o_work1.p, 1, 0, gs_op::add, 0, nullptr, st))) return rc;
  // (masked nodes carry id 0 in the masked handle, i.e. belong to no on-rank or halo row: ellipticApplyMask commutes
  // with the gather-scatter and rides along in its launch.  The reference splits fusedFDM into halo / interior
  // elements around oogs::start / finish, ellipticMultiGridSchwarz.cpp:1086-1121; with the one-launch exchange the
  // unsplit sequence has fewer launches and no fence.)
  if (options.compareArgs("MULTIGRID SMOOTHER", "RAS")) {
    if ((rc = fused_fdm_launch(mesh->Nq, 1, E, mesh->o_elementList.p, Su, o_Sx.p, o_Sy.p, o_Sz.p, o_invL.p,
                               elliptic->o_invDegreePfloat, o_work1.p, st)))
      return rc;
    return elliptic->oogs->startFinish<float>(Su, 1, 0, gs_op::add, elliptic->Nmasked, elliptic->o_maskIds.p, st);
  }
  // ASM
  if ((rc = fused_fdm_launch(mesh->Nq, 0, E, mesh->o_elementList.p, o_work2.p, o_Sx.p, o_Sy.p, o_Sz.p, o_invL.p,
                             nullptr, o_work1.p, st)))
    return rc;
  if ((rc = ogsExt->startFinish<float>(o_work2.p, 1, 0, gs_op::add, 0, nullptr, st))) return rc;
  if ((rc = post_fdm_launch(mesh->Nq, E, o_work1.p, o_work2.p, Su, o_wts.p, st))) return rc;
  return elliptic->oogs->startFinish<float>(Su, 1, 0, gs_op::add, elliptic->Nmasked, elliptic->o_maskIds.p, st);
}

// ------------------------------------------------------------------------------------------
// Schwarz setup (ellipticMultiGridSchwarz.cpp:58-1054)
// ------------------------------------------------------------------------------------------
namespace {
struct ElementLengths {
  std::vector<double> left[3], middle[3], right[3];
};

// harmonic_mean_element_length + compute_element_lengths, on the FINEST mesh (:953 "using the most
// refined level")
int compute_element_lengths(ElementLengths& L, elliptic_t* base)
{
  mesh_t* mesh = base->mesh;
  const dlong E = mesh->Nelements;
  const int N = mesh->N, Nq = mesh->Nq, Np = mesh->Np;
  const std::vector<double>& w = mesh->gllw;
  for (int d = 0; d < 3; ++d) {
    L.left[d].assign(E, 0.0);
    L.middle[d].assign(E, 0.0);
    L.right[d].assign(E, 0.0);
  }
  const int i1 = 0, i2 = Nq - 1;
  const int nx = (Nq == 2) ? Nq : Nq - 1;
  const int start = (Nq == 2) ? 0 : 1;
  const int strideOf[3] = {1, Nq, Nq * Nq};
  for (dlong e = 0; e < E; ++e) {
    const size_t off = (size_t)e * Np;
    for (int d = 0; d < 3; ++d) {
      // the two directions transverse to d, in the reference's loop order (outer, inner)
      const int so = (d == 2) ? strideOf[1] : strideOf[2];            // r: k outer; s: k outer; t: j outer
      const int si = (d == 0) ? strideOf[1] : strideOf[0];            // r: j inner; s: i inner; t: i inner
      double l2 = 0.0, wsum = 0.0;
      for (int o = start; o < nx; ++o)
        for (int in = start; in < nx; ++in) {
          const size_t a = off + (size_t)i2 * strideOf[d] + (size_t)o * so + (size_t)in * si;
          const size_t b = off + (size_t)i1 * strideOf[d] + (size_t)o * so + (size_t)in * si;
          const double weight = (Nq == 2) ? 1.0 : w[in - 1] * w[o - 1];
          const double dx = mesh->x[a] - mesh->x[b], dy = mesh->y[a] - mesh->y[b], dz = mesh->z[a] - mesh->z[b];
          l2 += weight / (dx * dx + dy * dy + dz * dz);
          wsum += weight;
        }
      l2 /= wsum;
      L.middle[d][e] = 1.0 / std::sqrt(l2);
    }
  }
  const double tol = 1e-12;
  for (dlong e = 0; e < E; ++e)
    for (int d = 0; d < 3; ++d) {
      const double m = L.middle[d][e];
      if (std::fabs(m) < tol || m < -tol || std::isnan(m) || std::isinf(m)) {
        set_last_error("FDM setup: element with zero, negative or invalid length");
        return NRSB_ERR_INVALID;
      }
    }
  if (Nq == 2) {
    for (int d = 0; d < 3; ++d) {
      L.left[d] = L.middle[d];
      L.right[d] = L.middle[d];
    }
    return NRSB_OK;
  }
  std::vector<double> l((size_t)Np * E, 0.0);
  for (dlong e = 0; e < E; ++e) {
    const size_t off = (size_t)Np * e;
    for (int j = 1; j < N; ++j)
      for (int k = 1; k < N; ++k) {
        l[k * Nq + j * Nq * Nq + off] = L.middle[0][e];
        l[Nq - 1 + k * Nq + j * Nq * Nq + off] = L.middle[0][e];
        l[k + 0 * Nq + j * Nq * Nq + off] = L.middle[1][e];
        l[k + (Nq - 1) * Nq + j * Nq * Nq + off] = L.middle[1][e];
        l[k + j * Nq + off] = L.middle[2][e];
        l[k + j * Nq + (Nq - 1) * Nq * Nq + off] = L.middle[2][e];
      }
  }
  {
    dbuf<double> d;
    int rc;
    if ((rc = d.upload(l))) return rc;
    if ((rc = mesh->oogs->startFinish<double>(d.p, 1, 0, gs_op::add, 0, nullptr, nullptr))) return rc;
    NRSB_CUDA(cudaDeviceSynchronize());
    if ((rc = d.download(l))) return rc;
  }
  for (dlong e = 0; e < E; ++e) {
    const size_t off = (size_t)e * Np;
    L.left[0][e] = l[1 * Nq + 1 * Nq * Nq + off] - L.middle[0][e];
    L.right[0][e] = l[Nq - 1 + 1 * Nq + 1 * Nq * Nq + off] - L.middle[0][e];
    L.left[1][e] = l[1 + 1 * Nq * Nq + off] - L.middle[1][e];
    L.right[1][e] = l[1 + (Nq - 1) * Nq + 1 * Nq * Nq + off] - L.middle[1][e];
    L.left[2][e] = l[1 + Nq + off] - L.middle[2][e];
    L.right[2][e] = l[1 + Nq + (Nq - 1) * Nq * Nq + off] - L.middle[2][e];
  }
  for (dlong e = 0; e < E; ++e)
    for (int d = 0; d < 3; ++d) {
      if (std::fabs(L.left[d][e]) < tol || L.left[d][e] < -tol) L.left[d][e] = L.middle[d][e];
      if (std::fabs(L.right[d][e]) < tol || L.right[d][e] < -tol) L.right[d][e] = L.middle[d][e];
    }
  return NRSB_OK;
}

// compute_1d_stiffness_matrix / compute_1d_mass_matrix / compute_1d_matrices (:282-556)
int compute_1d_matrices(std::vector<double>& S, std::vector<double>& lam, int lbc, int rbc, double ll, double lm,
                        double lr, const mesh_t* mesh)
{
  const int n = mesh->N;
  const int nl = n + 3;
  const std::vector<double>& D = mesh->D;
  const std::vector<double>& gw = mesh->gllw;
  std::vector<double> ah((n + 1) * (n + 1)), tmp((n + 1) * (n + 1));
  for (int i = 0; i < n + 1; ++i)
    for (int j = 0; j < n + 1; ++j) tmp[i * (n + 1) + j] = D[i * (n + 1) + j] * gw[i];
  for (int i = 0; i < n + 1; ++i)
    for (int j = 0; j < n + 1; ++j) {
      double aij = 0.0;
      for (int k = 0; k < n + 1; ++k) aij += D[k * (n + 1) + i] * tmp[k * (n + 1) + j];
      ah[i + j * (n + 1)] = aij;
    }
  auto AH = [&](int i, int j) { return ah[i + (n + 1) * j]; };
  std::vector<double> a(nl * nl, 0.0), b(nl * nl, 0.0);
  auto A = [&](int i, int j) -> double& { return a[i + nl * j]; };
  auto B = [&](int i, int j) -> double& { return b[i + nl * j]; };
  const int i0 = (lbc == 1) ? 1 : 0;
  const int i1 = (rbc == 1) ? n - 1 : n;
  double fac = 2.0 / lm;
  A(1, 1) = 1.0;
  A(n + 1, n + 1) = 1.0;
  for (int j = i0; j <= i1; ++j)
    for (int i = i0; i <= i1; ++i) A(i + 1, j + 1) = fac * AH(i, j);
  if (lbc == 0) {
    fac = 2.0 / ll;
    A(0, 0) = fac * AH(n - 1, n - 1);
    A(1, 0) = fac * AH(n, n - 1);
    A(0, 1) = fac * AH(n - 1, n);
    A(1, 1) = A(1, 1) + fac * AH(n, n);
  } else {
    A(0, 0) = 1.0;
  }
  if (rbc == 0) {
    fac = 2.0 / lr;
    A(n + 1, n + 1) = A(n + 1, n + 1) + fac * AH(0, 0);
    A(n + 2, n + 1) = fac * AH(1, 0);
    A(n + 1, n + 2) = fac * AH(0, 1);
    A(n + 2, n + 2) = fac * AH(1, 1);
  } else {
    A(n + 2, n + 2) = 1.0;
  }
  fac = 0.5 * lm;
  B(1, 1) = 1.0;
  B(n + 1, n + 1) = 1.0;
  for (int i = i0; i <= i1; ++i) B(i + 1, i + 1) = fac * gw[i];
  if (lbc == 0) {
    fac = 0.5 * ll;
    B(0, 0) = fac * gw[n - 1];
    B(1, 1) = B(1, 1) + fac * gw[n];
  } else {
    B(0, 0) = 1.0;
  }
  if (rbc == 0) {
    fac = 0.5 * lr;
    B(n + 1, n + 1) = B(n + 1, n + 1) + fac * gw[0];
    B(n + 2, n + 2) = fac * gw[1];
  } else {
    B(n + 2, n + 2) = 1.0;
  }
  if (sym_generalized_eig(nl, a, b, lam)) {
    set_last_error("FDM setup: generalized eigenproblem failed (B not positive definite)");
    return NRSB_ERR_INVALID;
  }
  S = a;
  auto row_zero = [&](int offset) {
    for (int i = 0; i < nl; ++i) S[offset + nl * i] = 0.0;
  };
  if (lbc > 0) row_zero(0);
  if (lbc == 1) row_zero(1);
  if (rbc > 0) row_zero(nl - 1);
  if (rbc == 1) row_zero(nl - 2);
  return NRSB_OK;
}

// global ids of the extended (N+2) mesh restricted to what the Schwarz exchange needs: the
// (N+1)^2 interior nodes of every extended face, paired with the matching nodes of the face
// neighbour.  Everything else (extended edges/corners, element interiors) is 0 = not exchanged:
// the reference masks edges/corners (create_extended_mesh :707-737) and interiors are singletons.
void extended_face_ids(const mesh_t* mesh, std::vector<hlong>& ext)
{
  const int N = mesh->N, Nq = mesh->Nq, Np = mesh->Np, Nqe = Nq + 2;
  const size_t Npe = (size_t)Nqe * Nqe * Nqe;
  ext.assign(Npe * mesh->Nelements, 0);
  std::vector<std::pair<hlong, int>> face(Nq * Nq);
  for (dlong e = 0; e < mesh->Nelements; ++e) {
    const hlong* g = &mesh->globalIds[(size_t)e * Np];
    for (int f = 0; f < 6; ++f) {
      // (a,b) on the face -> element node and extended node
      auto nodeOf = [&](int a, int b, int& n, size_t& ne) {
        int i, j, k, ie, je, ke;
        switch (f) {
          case 0: i = a; j = b; k = 0; ie = a + 1; je = b + 1; ke = 0; break;
          case 5: i = a; j = b; k = N; ie = a + 1; je = b + 1; ke = Nqe - 1; break;
          case 1: i = a; j = 0; k = b; ie = a + 1; je = 0; ke = b + 1; break;
          case 3: i = a; j = N; k = b; ie = a + 1; je = Nqe - 1; ke = b + 1; break;
          case 4: i = 0; j = a; k = b; ie = 0; je = a + 1; ke = b + 1; break;
          default: i = N; j = a; k = b; ie = Nqe - 1; je = a + 1; ke = b + 1; break;
        }
        n = i + Nq * j + Nq * Nq * k;
        ne = (size_t)ie + (size_t)Nqe * je + (size_t)Nqe * Nqe * ke;
      };
      hlong key;
      if (N >= 2) {
        int n;
        size_t ne;
        // smallest id among the face-interior nodes identifies the geometric face
        key = -1;
        for (int b = 1; b < N; ++b)
          for (int a = 1; a < N; ++a) {
            nodeOf(a, b, n, ne);
            if (key < 0 || g[n] < key) key = g[n];
          }
      } else {
        hlong c[4];
        int n;
        size_t ne;
        int q = 0;
        for (int b = 0; b < 2; ++b)
          for (int a = 0; a < 2; ++a) {
            nodeOf(a, b, n, ne);
            c[q++] = g[n];
          }
        std::sort(c, c + 4);
        uint64_t h = splitmix64((uint64_t)c[0]);
        for (int t = 1; t < 4; ++t) h = splitmix64(h ^ (uint64_t)c[t]);
        key = (hlong)(h >> 10);
      }
      for (int b = 0; b < Nq; ++b)
        for (int a = 0; a < Nq; ++a) {
          int n;
          size_t ne;
          nodeOf(a, b, n, ne);
          face[a + Nq * b] = {g[n], a + Nq * b};
        }
      std::sort(face.begin(), face.end());
      for (int r = 0; r < Nq * Nq; ++r) {
        int n;
        size_t ne;
        nodeOf(face[r].second % Nq, face[r].second / Nq, n, ne);
        ext[(size_t)e * Npe + ne] = key * 256 + r + 1;
      }
    }
  }
}
}  // namespace

int pMGLevel::generate_weights()
{
  // generate_weights (:798-838): count, through the two exchanges, how many Schwarz patches cover a node
  const dlong E = mesh->Nelements;
  const int Nq = mesh->Nq, Nqe = Nq + 2;
  const size_t weightSize = (size_t)Nq * Nq * Nq * E, extendedSize = (size_t)Nqe * Nqe * Nqe * E;
  std::vector<float> wts(weightSize), work1(extendedSize, 1.0f), work2(extendedSize, 1.0f);
  auto at = [&](std::vector<float>& v, int r, int s, int t, dlong e) -> float& {
    return v[(size_t)r + Nqe * ((size_t)s + Nqe * ((size_t)t + (size_t)Nqe * e))];
  };
  auto extrude = [&](std::vector<float>& a1, int l1, float f1, std::vector<float>& a2, int l2, float f2) {
    const int i0 = 1, i1 = Nqe - 1;
    for (dlong ie = 0; ie < E; ++ie) {
      for (int k = i0; k < i1; ++k)
        for (int j = i0; j < i1; ++j) {
          at(a1, l1, j, k, ie) = f1 * at(a1, l1, j, k, ie) + f2 * at(a2, l2, j, k, ie);
          at(a1, Nqe - l1 - 1, j, k, ie) = f1 * at(a1, Nqe - l1 - 1, j, k, ie) + f2 * at(a2, Nqe - l2 - 1, j, k, ie);
        }
      for (int k = i0; k < i1; ++k)
        for (int i = i0; i < i1; ++i) {
          at(a1, i, l1, k, ie) = f1 * at(a1, i, l1, k, ie) + f2 * at(a2, i, l2, k, ie);
          at(a1, i, Nqe - l1 - 1, k, ie) = f1 * at(a1, i, Nqe - l1 - 1, k, ie) + f2 * at(a2, i, Nqe - l2 - 1, k, ie);
        }
      for (int j = i0; j < i1; ++j)
        for (int i = i0; i < i1; ++i) {
          at(a1, i, j, l1, ie) = f1 * at(a1, i, j, l1, ie) + f2 * at(a2, i, j, l2, ie);
          at(a1, i, j, Nqe - l1 - 1, ie) = f1 * at(a1, i, j, Nqe - l1 - 1, ie) + f2 * at(a2, i, j, Nqe - l2 - 1, ie);
        }
    }
  };
  int rc;
  extrude(work2, 0, 0.0f, work1, 0, 1.0f);
  {
    dbuf<float> d;
    if ((rc = d.upload(work1))) return rc;
    if ((rc = ogsExt->startFinish<float>(d.p, 1, 0, gs_op::add, 0, nullptr, nullptr))) return rc;
    NRSB_CUDA(cudaDeviceSynchronize());
    if ((rc = d.download(work1))) return rc;
  }
  extrude(work1, 0, 1.0f, work2, 0, -1.0f);
  extrude(work1, 2, 1.0f, work1, 0, 1.0f);
  for (dlong ie = 0; ie < E; ++ie)
    for (int k = 0; k < Nq; ++k)
      for (int j = 0; j < Nq; ++j)
        for (int i = 0; i < Nq; ++i)
          wts[(size_t)i + Nq * ((size_t)j + Nq * ((size_t)k + (size_t)Nq * ie))] = at(work1, i + 1, j + 1, k + 1, ie);
  {
    dbuf<float> d;
    if ((rc = d.upload(wts))) return rc;
    if ((rc = elliptic->oogs->startFinish<float>(d.p, 1, 0, gs_op::add, 0, nullptr, nullptr))) return rc;
    NRSB_CUDA(cudaDeviceSynchronize());
    if ((rc = d.download(wts))) return rc;
  }
  for (auto& v : wts) v = 1.0f / v;
  return o_wts.upload(wts);
}

int pMGLevel::buildSchwarz()
{
  const dlong E = mesh->Nelements;
  const int Nq = mesh->Nq, Nqe = Nq + 2, Npe = Nqe * Nqe * Nqe;
  int rc;
  // extended-mesh exchange
  std::vector<hlong> extIds;
  extended_face_ids(mesh, extIds);
  ogsExtData.reset(new ogs_t());
  // which extended ids are shared with other ranks: derived from the base topology of the level mesh by the
  // same face construction on the other side, i.e. an id is shared iff its face key node is shared.
  SharedTopology extTopo;
  std::vector<hlong> sIds;
  std::vector<int> sOff, sRanks;
  const SharedTopology* tp = nullptr;
  if (mesh->topo.nranks > 1) {
    // a face-interior extended id is shared with exactly the ranks sharing ALL nodes of the face;
    // for a conforming mesh that is the sharer set of the order-N face-centre key node.  For N = 1 the
    // key is a hash of the four corners, so the sharer set is the intersection of the corners' sets.
    std::map<hlong, std::vector<int>> sharers;
    auto sharersOf = [&](hlong gid, std::vector<int>& out) {
      out.clear();
      const hlong* b = mesh->topo.sharedIds;
      const hlong* e_ = b + mesh->topo.nShared;
      const hlong* p = std::lower_bound(b, e_, gid);
      if (p == e_ || *p != gid) return;
      const long s = p - b;
      out.assign(mesh->topo.sharerRanks + mesh->topo.sharerOffsets[s],
                 mesh->topo.sharerRanks + mesh->topo.sharerOffsets[s + 1]);
    };
    const int N = mesh->N, Np = mesh->Np;
    for (dlong e = 0; e < E; ++e) {
      const hlong* g = &mesh->globalIds[(size_t)e * Np];
      for (size_t ne = 0; ne < (size_t)Npe; ++ne) {
        const hlong id = extIds[(size_t)e * Npe + ne];
        if (id == 0 || sharers.count(id)) continue;
        // recover the face from the extended node position
        const int ie = ne % Nqe, je = (ne / Nqe) % Nqe, ke = ne / (Nqe * Nqe);
        int corner[4];
        auto nid = [&](int i, int j, int k) { return i + Nq * j + Nq * Nq * k; };
        if (ke == 0 || ke == Nqe - 1) {
          const int k = ke ? N : 0;
          corner[0] = nid(0, 0, k); corner[1] = nid(N, 0, k); corner[2] = nid(0, N, k); corner[3] = nid(N, N, k);
        } else if (je == 0 || je == Nqe - 1) {
          const int j = je ? N : 0;
          corner[0] = nid(0, j, 0); corner[1] = nid(N, j, 0); corner[2] = nid(0, j, N); corner[3] = nid(N, j, N);
        } else {
          const int i = ie ? N : 0;
          corner[0] = nid(i, 0, 0); corner[1] = nid(i, N, 0); corner[2] = nid(i, 0, N); corner[3] = nid(i, N, N);
        }
        std::vector<int> acc, cur, tmp;
        sharersOf(g[corner[0]], acc);
        for (int c = 1; c < 4 && !acc.empty(); ++c) {
          sharersOf(g[corner[c]], cur);
          tmp.clear();
          std::set_intersection(acc.begin(), acc.end(), cur.begin(), cur.end(), std::back_inserter(tmp));
          acc.swap(tmp);
        }
        if (acc.size() > 1) sharers[id] = acc;
      }
    }
    sOff.push_back(0);
    for (auto& kv : sharers) {
      sIds.push_back(kv.first);
      for (int r : kv.second) sRanks.push_back(r);
      sOff.push_back((int)sRanks.size());
    }
    extTopo.rank = mesh->topo.rank;
    extTopo.nranks = mesh->topo.nranks;
    extTopo.nShared = (long)sIds.size();
    extTopo.sharedIds = sIds.data();
    extTopo.sharerOffsets = sOff.data();
    extTopo.sharerRanks = sRanks.data();
    tp = &extTopo;
  }
  if ((rc = ogsExtData->setup((dlong)((size_t)E * Npe), extIds.data(), tp))) return rc;
  ogsExt.reset(new oogs_t());
  if ((rc = ogsExt->setup(ogsExtData.get(), mesh->comm, 1))) return rc;

  // element lengths from the finest level, 1-D operators, eigen-decompositions
  ElementLengths L;
  if ((rc = compute_element_lengths(L, ellipticBase))) return rc;
  std::vector<float> Sx((size_t)Nqe * Nqe * E), Sy(Sx.size()), Sz(Sx.size()), invL((size_t)Npe * E);
  std::vector<double> S[3], lam[3];
  const int lookup[] = {4, 2, 1, 3, 0, 5};  // compute_element_boundary_conditions (:262-280)
  for (dlong e = 0; e < E; ++e) {
    int fbc[6];
    for (int iface = 0; iface < 6; ++iface) fbc[iface] = elliptic->EToB[6 * e + lookup[iface]];
    for (int d = 0; d < 3; ++d)
      if ((rc = compute_1d_matrices(S[d], lam[d], fbc[2 * d], fbc[2 * d + 1], L.left[d][e], L.middle[d][e],
                                    L.right[d][e], mesh)))
        return rc;
    // "store the transposes" (:607-616): row-major [node][mode]
    for (int i = 0; i < Nqe; ++i)
      for (int j = 0; j < Nqe; ++j) {
        const size_t o = (size_t)Nqe * Nqe * e + i + j * Nqe;
        Sx[o] = (float)S[0][j + i * Nqe];
        Sy[o] = (float)S[1][j + i * Nqe];
        Sz[o] = (float)S[2][j + i * Nqe];
      }
    size_t l = 0;
    for (int k = 0; k < Nqe; ++k)
      for (int j = 0; j < Nqe; ++j)
        for (int i = 0; i < Nqe; ++i) {
          const double diag = lam[0][i] + lam[1][j] + lam[2][k];
          invL[(size_t)Npe * e + l] = (float)((diag > 1e-5) ? 1.0 / diag : 0.0);
          ++l;
        }
  }
  if ((rc = o_Sx.upload(Sx))) return rc;
  if ((rc = o_Sy.upload(Sy))) return rc;
  if ((rc = o_Sz.upload(Sz))) return rc;
  if ((rc = o_invL.upload(invL))) return rc;
  if ((rc = o_work1.alloc((size_t)Npe * E))) return rc;
  if (!options.compareArgs("MULTIGRID SMOOTHER", "RAS"))
    if ((rc = o_work2.alloc((size_t)Npe * E))) return rc;
  return generate_weights();
}

// ------------------------------------------------------------------------------------------
// maxEigSmoothAx: Arnoldi(10) on S*A (ellipticMultiGridLevelSetup.cpp:292-453).  The reference starts
// from std::random_device noise (randomVector.hpp:15-23, not reproducible); here the start vector
// is a fixed hash of the global node id so that runs, rank counts and the oracle all agree.
// ------------------------------------------------------------------------------------------
int pMGLevel::maxEigSmoothAx(double* rhoOut)
{
  cudaStream_t st = elliptic->stream;
  const long M = Nrows;
  hlong Nglobal = (hlong)mesh->NelementsGlobal * mesh->Np;
  const int k = (int)std::min<hlong>(10, Nglobal);
  std::vector<double> H((size_t)k * k, 0.0);
  std::vector<dbuf<double>> V(k + 1);
  int rc;
  for (auto& v : V)
    if ((rc = v.alloc(M))) return rc;
  dbuf<double> o_Vx;
  dbuf<float> o_VxPfloat, o_AVxPfloat;
  if ((rc = o_VxPfloat.alloc(M))) return rc;
  if ((rc = o_AVxPfloat.alloc(M))) return rc;
  std::vector<double> Vx(M);
  for (long n = 0; n < M; ++n) Vx[n] = id_uniform(mesh->globalIds[n]);
  if ((rc = o_Vx.upload(Vx))) return rc;
  // gs over the unmasked mesh handle, then zero the Dirichlet rows
  if ((rc = mesh->oogs->startFinish<double>(o_Vx.p, 1, 0, gs_op::add, 0, nullptr, st))) return rc;
  if ((rc = ellipticApplyMask<double>(elliptic, o_Vx.p))) return rc;
  double* S = elliptic->o_scal.p;
  const double* w = elliptic->ogs->d_invDegree;  // elliptic->ogs->invDegree (:305)
  auto dot = [&](const double* a, const double* b, double* out) -> int {
    int r = wdot_launch<double>(M, w, a, b, S, elliptic->ws, st);
    if (r) return r;
    return elliptic->read_scalars(0, 1, out);
  };
  double norm_vo;
  if ((rc = dot(o_Vx.p, o_Vx.p, &norm_vo))) return rc;
  norm_vo = std::sqrt(norm_vo);
  if ((rc = axpby_launch<double>(M, DevScalar::host(1. / norm_vo), o_Vx.p, DevScalar::host(0.0), V[0].p, st)))
    return rc;
  for (int j = 0; j < k; j++) {
    if ((rc = copy_d2f_launch(M, V[j].p, o_VxPfloat.p, st))) return rc;
    if ((rc = ellipticOperator<float>(elliptic, o_VxPfloat.p, o_AVxPfloat.p))) return rc;
    if ((rc = smoother(o_AVxPfloat.p, o_VxPfloat.p, true))) return rc;
    if ((rc = copy_f2d_launch(M, o_VxPfloat.p, V[j + 1].p, st))) return rc;
    for (int i = 0; i <= j; i++) {
      double hij;
      if ((rc = dot(V[i].p, V[j + 1].p, &hij))) return rc;
      if ((rc = axpby_launch<double>(M, DevScalar::host(-hij), V[i].p, DevScalar::host(1.0), V[j + 1].p, st)))
        return rc;
      H[i + j * k] = hij;
    }
    if (j + 1 < k) {
      double norm_vj;
      if ((rc = dot(V[j + 1].p, V[j + 1].p, &norm_vj))) return rc;
      norm_vj = std::sqrt(norm_vj);
      if ((rc = scale_launch<double>(M, 1 / norm_vj, V[j + 1].p, st))) return rc;
      H[j + 1 + j * k] = norm_vj;
    }
  }
  for (double v : H)
    if (std::isnan(v) || std::isinf(v)) {
      set_last_error("maxEigSmoothAx: invalid matrix entries!");
      return NRSB_ERR_DIVERGED;
    }
  *rhoOut = hessenberg_spectral_radius(k, H);
  return NRSB_OK;
}

// setupSmoother (ellipticMultiGridLevelSetup.cpp:109-180)
int pMGLevel::setupSmoother()
{
  double minMultiplier = 0.1, maxMultiplier = 1.1;
  options.getArgs("MULTIGRID CHEBYSHEV MIN EIGENVALUE BOUND FACTOR", minMultiplier);
  options.getArgs("MULTIGRID CHEBYSHEV MAX EIGENVALUE BOUND FACTOR", maxMultiplier);
  const bool useASM = options.compareArgs("MULTIGRID SMOOTHER", "ASM");
  const bool useRAS = options.compareArgs("MULTIGRID SMOOTHER", "RAS");
  const bool useJacobi = options.compareArgs("MULTIGRID SMOOTHER", "DAMPEDJACOBI");
  int rc;
  if ((rc = o_smootherResidual.alloc(Nrows))) return rc;
  if ((rc = o_smootherResidual2.alloc(Nrows))) return rc;
  if ((rc = o_smootherUpdate.alloc(Nrows))) return rc;
  if (useASM || useRAS) {
    smootherType = useASM ? SmootherType::ASM : SmootherType::RAS;
    if ((rc = buildSchwarz())) return rc;
  } else {
    NRSB_REQUIRE(useJacobi, "Invalid pMGLevel smoother!");
    smootherType = SmootherType::JACOBI;
    if ((rc = o_invDiagA.alloc(mesh->Nlocal))) return rc;
    if ((rc = ellipticBuildDiagonal<float>(elliptic, o_invDiagA.p))) return rc;
  }
  hasSmoother = true;
  if (options.compareArgs("MULTIGRID SMOOTHER", "CHEBYSHEV")) {
    chebySmootherType = smootherType == SmootherType::ASM
                            ? ChebyshevSmootherType::ASM
                            : (smootherType == SmootherType::RAS ? ChebyshevSmootherType::RAS
                                                                 : ChebyshevSmootherType::JACOBI);
    smootherType = SmootherType::CHEBYSHEV;
    double rho;
    if ((rc = maxEigSmoothAx(&rho))) return rc;
    lambda1 = maxMultiplier * rho;
    lambda0 = minMultiplier * rho;
    maxEig = rho;
    UpLegChebyshevDegree = 3;
    DownLegChebyshevDegree = 3;
    if (!isCoarse) {
      options.getArgs("MULTIGRID CHEBYSHEV DEGREE", UpLegChebyshevDegree);
      options.getArgs("MULTIGRID CHEBYSHEV DEGREE", DownLegChebyshevDegree);
    }
  }
  const std::string schedule = options.getArgs("MULTIGRID SCHEDULE");
  if (!schedule.empty()) {
    const int up = schedule_degree(schedule, degree, false), down = schedule_degree(schedule, degree, true);
    if (up > -1) UpLegChebyshevDegree = up;
    if (down > -1) DownLegChebyshevDegree = down;
  }
  if (options.compareArgs("MULTIGRID SMOOTHER", "FOURTHOPT")) {
    UpLegBetas = optimal_coeffs(UpLegChebyshevDegree);
    DownLegBetas = optimal_coeffs(DownLegChebyshevDegree);
    smootherType = SmootherType::OPT_FOURTH_CHEBYSHEV;
  } else if (options.compareArgs("MULTIGRID SMOOTHER", "FOURTH")) {
    UpLegBetas.assign(UpLegChebyshevDegree, 1.0f);
    DownLegBetas.assign(DownLegChebyshevDegree, 1.0f);
    smootherType = SmootherType::FOURTH_CHEBYSHEV;
  }
  return NRSB_OK;
}

// ------------------------------------------------------------------------------------------
// V-cycle (MGSolver.cpp:150-193)
// ------------------------------------------------------------------------------------------
int MGSolver_t::Run(float* o_rhs, float* o_x)
{
  levels[0]->o_x = o_x;
  levels[0]->o_rhs = o_rhs;
  return additive ? runAdditiveVcycle() : runVcycle(0);
}

// MGSolver.cpp:195-251 (with coarsenV, schwarzSolve, prolongateV of :32-78).  The reference overlaps the CPU
// (hypre) coarse solve with the device smoothing in a second OpenMP task; here everything is on the device and
// the coarse solve simply follows on the stream.
int MGSolver_t::runAdditiveVcycle()
{
  const int numLevels = (int)levels.size();
  int rc;
  cudaStream_t st = levels[0]->elliptic->stream;
  // coarsenV: rhs of every level = restriction of the level above
  for (int k = 0; k < numLevels - 1; ++k) {
    pMGLevel *level = levels[k].get(), *levelC = levels[k + 1].get();
    NRSB_CUDA(cudaMemcpyAsync(level->o_res.p, level->o_rhs, sizeof(float) * level->Nrows, cudaMemcpyDeviceToDevice, st));
    if ((rc = levelC->coarsen(level->o_res.p, levelC->o_rhs))) return rc;
  }
  // schwarzSolve: x_k = S_k rhs_k on every level but the coarsest (the reference repeats the restriction here)
  for (int k = 0; k < numLevels - 1; ++k) {
    pMGLevel *level = levels[k].get(), *levelC = levels[k + 1].get();
    if ((rc = level->smooth(level->o_rhs, level->o_x, true))) return rc;
    NRSB_CUDA(cudaMemcpyAsync(level->o_res.p, level->o_rhs, sizeof(float) * level->Nrows, cudaMemcpyDeviceToDevice, st));
    if ((rc = levelC->coarsen(level->o_res.p, levelC->o_rhs))) return rc;
  }
  pMGLevel* base = levels[baseLevel].get();
  if ((rc = coarseSolve(base->o_rhs, base->o_x))) return rc;
  // prolongateV: x_k += P x_{k+1}
  for (int k = numLevels - 2; k >= 0; --k)
    if ((rc = levels[k + 1]->prolongate(levels[k + 1]->o_x, levels[k]->o_x))) return rc;
  return NRSB_OK;
}

// developer aid (NRSB_MG_TIMING=1): CUDA-event time per V-cycle phase and level, printed every 32 cycles
namespace {
struct MgTimer {
  bool on = getenv("NRSB_MG_TIMING") != nullptr;
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> label;
  std::map<std::string, double> acc;
  int cycles = 0;
  void mark(const std::string& what, cudaStream_t st)
  {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    label.push_back(what);
  }
  void flush(int rank)
  {
    if (!on || ev.empty()) return;
    cudaEventSynchronize(ev.back());
    for (size_t i = 1; i < ev.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      acc[label[i]] += ms;
    }
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
    label.clear();
    if (++cycles % 32 == 0) {
      fprintf(stderr, "[rank %d] V-cycle phases, mean us per cycle over %d cycles:", rank, cycles);
      for (auto& kv : acc) fprintf(stderr, "  %s %.1f", kv.first.c_str(), kv.second / cycles * 1e3);
      fprintf(stderr, "\n");
    }
  }
};
MgTimer g_mgTimer;
}  // namespace

int MGSolver_t::runVcycle(int k)
{
  pMGLevel* level = levels[k].get();
  float *o_rhs = level->o_rhs, *o_x = level->o_x, *o_res = level->o_res.p;
  cudaStream_t st = level->elliptic->stream;
  const std::string L = "L" + std::to_string(k) + ":";
  if (k == 0) g_mgTimer.mark("start", st);
  if (k == baseLevel) {
    const int rc = coarseSolve(o_rhs, o_x);
    g_mgTimer.mark("coarse", st);
    return rc;
  }
  pMGLevel* levelC = levels[k + 1].get();
  int rc;
  if ((rc = level->smooth(o_rhs, o_x, true))) return rc;
  g_mgTimer.mark(L + "smoothDown", st);
  if ((rc = level->residual(o_rhs, o_x, o_res))) return rc;
  if ((rc = levelC->coarsen(o_res, levelC->o_rhs))) return rc;
  g_mgTimer.mark(L + "resid+coarsen", st);
  if ((rc = runVcycle(k + 1))) return rc;
  if ((rc = levelC->prolongate(levelC->o_x, o_x))) return rc;
  g_mgTimer.mark(L + "prolong", st);
  rc = level->smooth(o_rhs, o_x, false);
  g_mgTimer.mark(L + "smoothUp", st);
  if (k == 0) g_mgTimer.flush(level->elliptic->comm ? level->elliptic->comm->rank : 0);
  return rc;
}

// ------------------------------------------------------------------------------------------
// level construction (ellipticMultiGridSetup.cpp, ellipticBuildMultigridLevel*.cpp, createMeshMG)
// ------------------------------------------------------------------------------------------
static int build_level_elliptic(elliptic_t* base, mesh_t* baseMesh, int Nc, std::unique_ptr<mesh_t>& meshOut,
                                std::unique_ptr<elliptic_t>& ellOut)
{
  int rc;
  mesh_t* mesh = baseMesh;
  if (Nc != baseMesh->N) {
    // createMeshMG: interpolate the finest nodes to the order-Nc GLL points (map_m_to_n), new global
    // numbering of the order-Nc mesh, geometric factors
    meshOut.reset(new mesh_t());
    std::vector<double> zc, wc, I;
    mesh_t::gll(Nc, zc, wc);
    mesh_t::interp_matrix(baseMesh->gllz, zc, I);  // [Nqc][Nqf]
    const int Nqf = baseMesh->Nq, Nqc = Nc + 1;
    const dlong E = baseMesh->Nelements;
    std::vector<double> xc((size_t)E * Nqc * Nqc * Nqc), yc(xc.size()), zcn(xc.size());
    std::vector<double> t1((size_t)Nqf * Nqf * Nqc), t2((size_t)Nqf * Nqc * Nqc);
    auto interp = [&](const std::vector<double>& src, std::vector<double>& dst) {
      for (dlong e = 0; e < E; ++e) {
        const double* s = &src[(size_t)e * Nqf * Nqf * Nqf];
        for (int k = 0; k < Nqf; ++k)
          for (int j = 0; j < Nqf; ++j)
            for (int a = 0; a < Nqc; ++a) {
              double v = 0;
              for (int i = 0; i < Nqf; ++i) v += I[(size_t)a * Nqf + i] * s[i + Nqf * j + Nqf * Nqf * k];
              t1[a + Nqc * (j + Nqf * k)] = v;
            }
        for (int k = 0; k < Nqf; ++k)
          for (int b = 0; b < Nqc; ++b)
            for (int a = 0; a < Nqc; ++a) {
              double v = 0;
              for (int j = 0; j < Nqf; ++j) v += I[(size_t)b * Nqf + j] * t1[a + Nqc * (j + Nqf * k)];
              t2[a + Nqc * (b + Nqc * k)] = v;
            }
        double* d = &dst[(size_t)e * Nqc * Nqc * Nqc];
        for (int c = 0; c < Nqc; ++c)
          for (int b = 0; b < Nqc; ++b)
            for (int a = 0; a < Nqc; ++a) {
              double v = 0;
              for (int k = 0; k < Nqf; ++k) v += I[(size_t)c * Nqf + k] * t2[a + Nqc * (b + Nqc * k)];
              d[a + Nqc * (b + Nqc * c)] = v;
            }
      }
    };
    interp(baseMesh->x, xc);
    interp(baseMesh->y, yc);
    interp(baseMesh->z, zcn);
    // coarse global numbering: supplied by the caller through elliptic->options-independent table
    auto it = base->levelGlobalIds.find(Nc);
    if (it == base->levelGlobalIds.end()) {
      set_last_error("no global numbering supplied for multigrid level N=" + std::to_string(Nc));
      return NRSB_ERR_INVALID;
    }
    const SharedTopology* tp = nullptr;
    auto tIt = base->levelTopology.find(Nc);
    if (tIt != base->levelTopology.end()) tp = &tIt->second;
    if ((rc = meshOut->setup(Nc, E, xc.data(), yc.data(), zcn.data(), it->second.data(), baseMesh->EToB.data(),
                             baseMesh->comm, tp, /*keepFp64Geo=*/Nc == 1)))
      return rc;
    mesh = meshOut.get();
  }
  ellOut.reset(new elliptic_t());
  elliptic_t* e = ellOut.get();
  e->name = base->name;
  e->options = base->options;
  e->mesh = mesh;
  e->comm = base->comm;
  e->mgLevel = true;
  e->poisson = base->poisson;
  e->allNeumann = base->allNeumann;
  e->lambda0Value = base->lambda0Value;
  e->lambda1Value = base->lambda1Value;
  e->EToB = base->EToB;
  e->fieldOffset = mesh->Nlocal;
  e->stream = base->stream;
  e->ax_variant[0] = e->ax_variant[1] = -1;
  if ((rc = elliptic_workspace(e))) return rc;
  std::vector<float> l0f(1, (float)e->lambda0Value), l1f(1, (float)e->lambda1Value);
  if ((rc = e->o_lambda0Pfloat.upload(l0f))) return rc;
  if ((rc = e->o_lambda1Pfloat.upload(l1f))) return rc;
  std::vector<double> l0(1, e->lambda0Value), l1(1, e->lambda1Value);
  if ((rc = e->o_lambda0.upload(l0))) return rc;
  if ((rc = e->o_lambda1.upload(l1))) return rc;
  if ((rc = ellipticOgs(mesh, e->EToB, e))) return rc;
  e->overlap = e->ogs->NhaloGather > 0 && !e->options.compareArgs("ENABLE GS COMM OVERLAP", "FALSE") &&
               mesh->NlocalGatherElements > 0;
  return ellipticChooseOverlap(e, 4);
}

int ellipticMultiGridSetup(elliptic_t* elliptic_, precon_t* precon)
{
  options_t& options = elliptic_->options;
  mesh_t* mesh = elliptic_->mesh;
  std::vector<int> levelDegree = determineMGLevels(options, mesh->N);
  NRSB_REQUIRE(!levelDegree.empty() && levelDegree[0] == mesh->N, "multigrid schedule must start at the solver order");
  const int numMGLevels = (int)levelDegree.size();
  const int Nmax = levelDegree[0], Nmin = levelDegree[numMGLevels - 1];
  // every transfer pair and Schwarz size of this schedule must be instantiated: fail HERE with a clear message,
  // not in the first V-cycle
  {
    const bool schwarzSm =
        options.compareArgs("MULTIGRID SMOOTHER", "ASM") || options.compareArgs("MULTIGRID SMOOTHER", "RAS");
    for (int n = 0; n < numMGLevels; ++n) {
      if (n > 0 && !transfer_supported(levelDegree[n - 1] + 1, levelDegree[n] + 1)) {
        set_last_error("multigrid schedule needs the coarsen/prolongate pair N=" + std::to_string(levelDegree[n - 1]) +
                       " -> N=" + std::to_string(levelDegree[n]) + ", which is not instantiated (transfer.cu)");
        return NRSB_ERR_INVALID;
      }
      if (schwarzSm && levelDegree[n] > 1 && !fdm_supported(levelDegree[n] + 1)) {
        set_last_error("Schwarz smoother at N=" + std::to_string(levelDegree[n]) + " is not instantiated (fdm.cu)");
        return NRSB_ERR_INVALID;
      }
    }
  }
  precon->MGSolver.reset(new MGSolver_t());
  MGSolver_t* mg = precon->MGSolver.get();
  // cycle type and its legality (MGSolver.cpp:99-139)
  {
    const bool cheb = options.compareArgs("MULTIGRID SMOOTHER", "CHEBYSHEV");
    const bool schwarzSm =
        options.compareArgs("MULTIGRID SMOOTHER", "ASM") || options.compareArgs("MULTIGRID SMOOTHER", "RAS");
    if (options.has("MGSOLVER CYCLE") && !options.compareArgs("MGSOLVER CYCLE", "VCYCLE")) {
      set_last_error("Unknown multigrid cycle type!");
      return NRSB_ERR_INVALID;
    }
    mg->additive = options.compareArgs("MGSOLVER CYCLE", "ADDITIVE");
    if (mg->additive && cheb) {
      set_last_error("Additive vcycle is not supported for Chebyshev!");
      return NRSB_ERR_INVALID;
    }
    if (!mg->additive && schwarzSm && !cheb) {
      set_last_error("Multiplicative vcycle is not supported for RAS/ASM smoother without Chebyshev!");
      return NRSB_ERR_INVALID;
    }
  }
  const bool coarseSolveOpt = options.compareArgs("MULTIGRID COARSE SOLVE", "TRUE");
  const bool coarseAndSmooth = options.compareArgs("MULTIGRID COARSE SOLVE AND SMOOTH", "TRUE");
  int rc;
  auto newLevel = [&](elliptic_t* e, int degree, bool isCoarse) {
    auto lvl = std::make_unique<pMGLevel>();
    lvl->elliptic = e;
    lvl->ellipticBase = elliptic_;
    lvl->mesh = e->mesh;
    lvl->options = options;
    lvl->degree = degree;
    lvl->isCoarse = isCoarse;
    lvl->Nrows = e->mesh->Nlocal;
    return lvl;
  };
  elliptic_t* fine = nullptr;
  for (int n = 0; n < numMGLevels; ++n) {
    const int Nc = levelDegree[n];
    const bool isCoarse = (n == numMGLevels - 1);
    std::unique_ptr<mesh_t> m;
    std::unique_ptr<elliptic_t> e;
    if ((rc = build_level_elliptic(elliptic_, mesh, Nc, m, e))) return rc;
    auto lvl = newLevel(e.get(), Nc, isCoarse);
    if (n > 0) {
      // buildCoarsenerQuadHex (:238-258): R = (coarse->fine interpolation)^T, [NqC][NqF]
      const int Nf = levelDegree[n - 1];
      std::vector<double> zf, wf, zc, wc, P;
      mesh_t::gll(Nf, zf, wf);
      mesh_t::gll(Nc, zc, wc);
      mesh_t::interp_matrix(zc, zf, P);  // [Nqf][Nqc]
      lvl->R.assign((size_t)(Nc + 1) * (Nf + 1), 0.f);
      for (int i = 0; i < Nc + 1; ++i)
        for (int j = 0; j < Nf + 1; ++j) lvl->R[(size_t)i * (Nf + 1) + j] = (float)P[(size_t)j * (Nc + 1) + i];
      lvl->NqF = Nf + 1;
      lvl->NpF = (dlong)(Nf + 1) * (Nf + 1) * (Nf + 1);
      // nodal interpolation fine -> coarse for the coefficient fields (ellipticBuildMultigridLevel.cpp:133-146)
      {
        std::vector<double> I;
        mesh_t::interp_matrix(zf, zc, I);  // [Nqc][Nqf]
        e->interpFromFine.assign(I.begin(), I.end());
      }
      lvl->o_invDegreeFine = fine->o_invDegreePfloat;
      if ((rc = lvl->x_store.alloc(lvl->Nrows))) return rc;
      if ((rc = lvl->rhs_store.alloc(lvl->Nrows))) return rc;
      lvl->o_x = lvl->x_store.p;
      lvl->o_rhs = lvl->rhs_store.p;
    }
    if ((rc = lvl->o_res.alloc(lvl->Nrows))) return rc;
    const bool needSmoother = !isCoarse || numMGLevels == 1 || !coarseSolveOpt || coarseAndSmooth;
    if (needSmoother)
      if ((rc = lvl->setupSmoother())) return rc;
    fine = e.get();
    if (m) mg->meshLevels.push_back(std::move(m));
    mg->ellipticLevels.push_back(std::move(e));
    mg->levels.push_back(std::move(lvl));
  }
  mg->baseLevel = numMGLevels - 1;
  pMGLevel* baseLevel = mg->levels.back().get();
  (void)Nmax;
  (void)Nmin;
  if (coarseSolveOpt) {
    NRSB_REQUIRE(!options.compareArgs("COARSE SOLVER", "BOOMERAMG") && !options.compareArgs("COARSE SOLVER", "AMGX"),
                 "COARSE SOLVER BOOMERAMG/AMGX are third-party libraries outside this path; use JPCG or SMOOTHER");
    int maxIt = 200;
    double ctol = 1e-1;  // one BoomerAMG V-cycle is an inexact solve too; 1e-1 keeps the BPS5 iteration count (28)
    options.getArgs("COARSE SOLVER MAXIMUM ITERATIONS", maxIt);
    options.getArgs("COARSE SOLVER TOLERANCE", ctol);
    precon->coarse.reset(new coarseSolver_t());
    if ((rc = precon->coarse->setup(baseLevel, maxIt, ctol))) return rc;
    coarseSolver_t* cs = precon->coarse.get();
    if (coarseAndSmooth) {
      mg->coarseSolve = [baseLevel, cs](float* rhs, float* x) -> int {
        int r;
        float* res = baseLevel->o_res.p;
        float* tmp = baseLevel->o_smootherUpdate.p;
        if ((r = baseLevel->smooth(rhs, x, true))) return r;
        if ((r = baseLevel->residual(rhs, x, res))) return r;
        if ((r = cs->solve(res, tmp))) return r;
        if ((r = axpby_launch<float>(baseLevel->Nrows, DevScalar::host(1.0), tmp, DevScalar::host(1.0), x,
                                     baseLevel->elliptic->stream)))
          return r;
        return baseLevel->smooth(rhs, x, false);
      };
    } else {
      mg->coarseSolve = [cs](float* rhs, float* x) -> int { return cs->solve(rhs, x); };
    }
  } else {
    mg->coarseSolve = [baseLevel](float* rhs, float* x) -> int { return baseLevel->smooth(rhs, x, true); };
  }
  return NRSB_OK;
}

}  // namespace nrsb

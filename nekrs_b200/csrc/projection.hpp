// projection.hpp -- residual projection onto previous solutions
// (SolutionProjection, ellipticSolutionProjection.cpp:44-288; kernels accumulate.okl,
// multiScaledAddwOffset.okl).  Nfields == 1.
#pragma once
#include "host.hpp"

namespace nrsb {

int accumulate_launch(long N, int m, long fieldOffset, const double* alpha, const double* x, double* y,
                      cudaStream_t s);
int multi_scaled_add_w_offset_launch(long N, int m, long destOffset, long fieldOffset, const double* alphas,
                                     double beta, double* x, cudaStream_t s);

class SolutionProjection {
 public:
  elliptic_t* elliptic;
  bool aconj;
  int maxNumVecsProjection, numTimeSteps;
  int numVecsProjection = 0, prevNumVecsProjection = 0, timestep = 0;
  long Nlocal, fieldOffset;
  std::vector<double> alpha;
  dbuf<double> o_alpha, o_xbar, o_xx, o_bb;

  SolutionProjection(elliptic_t* e, bool aconj_, int maxVecs, int nSteps)
      : elliptic(e), aconj(aconj_), maxNumVecsProjection(maxVecs), numTimeSteps(nSteps)
  {
    Nlocal = e->mesh->Nlocal;
    fieldOffset = e->fieldOffset;
    alpha.assign(maxVecs, 0.0);
  }
  int setup()
  {
    int rc;
    NRSB_REQUIRE(maxNumVecsProjection >= 1 && maxNumVecsProjection <= kMaxRed,
                 "RESIDUAL PROJECTION VECTORS must be in 1..16");
    if ((rc = o_alpha.alloc(maxNumVecsProjection))) return rc;
    if ((rc = o_xbar.alloc(fieldOffset))) return rc;
    if ((rc = o_xx.alloc((size_t)fieldOffset * maxNumVecsProjection))) return rc;
    return o_bb.alloc(aconj ? (size_t)fieldOffset : (size_t)fieldOffset * maxNumVecsProjection);
  }
  const double* weight() const { return elliptic->mesh->ogs->d_invDegree; }  // mesh->ogs->o_invDegree (:243)

  int matvec(double* o_Ax, long Ax_offset, double* o_x, long x_offset)
  {
    return ellipticOperator<double>(elliptic, o_x + x_offset * fieldOffset, o_Ax + Ax_offset * fieldOffset);
  }
  int multi_dot(const double* x, const double* y, int nv)
  {
    cudaStream_t st = elliptic->stream;
    int rc = wdot_multi_launch(Nlocal, nv, fieldOffset, weight(), x, y, o_alpha.p, elliptic->ws, st);
    if (rc) return rc;
    NRSB_CUDA(cudaMemcpyAsync(alpha.data(), o_alpha.p, sizeof(double) * nv, cudaMemcpyDeviceToHost, st));
    NRSB_CUDA(cudaStreamSynchronize(st));
    return NRSB_OK;
  }
  int updateProjectionSpace()
  {
    if (numVecsProjection <= 0) return NRSB_OK;
    cudaStream_t st = elliptic->stream;
    const int m = numVecsProjection;
    int rc;
    if ((rc = multi_dot(o_xx.p, o_bb.p + (aconj ? 0 : (long)(m - 1) * fieldOffset), m))) return rc;
    const double norm_orig = alpha[m - 1];
    if ((rc = multi_scaled_add_w_offset_launch(Nlocal, m, (long)(m - 1) * fieldOffset, fieldOffset, o_alpha.p, 1.0,
                                               o_xx.p, st)))
      return rc;
    if (!aconj)
      if ((rc = multi_scaled_add_w_offset_launch(Nlocal, m, (long)(m - 1) * fieldOffset, fieldOffset, o_alpha.p, 1.0,
                                                 o_bb.p, st)))
        return rc;
    double sumAlpha = 0;
    for (int k = 0; k < m - 1; ++k) sumAlpha += alpha[k] * alpha[k];
    double norm_new = std::sqrt(norm_orig - sumAlpha);
    const double tol = 1e-7;
    if (norm_new / norm_orig > tol) {
      const double scale = 1.0 / norm_new;
      if ((rc = scale_launch<double>(Nlocal, scale, o_xx.p + (size_t)fieldOffset * (m - 1), st))) return rc;
      if (!aconj)
        if ((rc = scale_launch<double>(Nlocal, scale, o_bb.p + (size_t)fieldOffset * (m - 1), st))) return rc;
    } else {
      numVecsProjection--;  // linearly dependent: discard
    }
    return NRSB_OK;
  }
  int pre(double* o_r)
  {
    ++timestep;
    if (timestep < numTimeSteps) return NRSB_OK;
    if (numVecsProjection <= 0) return NRSB_OK;
    prevNumVecsProjection = numVecsProjection;
    cudaStream_t st = elliptic->stream;
    int rc;
    if ((rc = multi_dot(o_xx.p, o_r, numVecsProjection))) return rc;
    if ((rc = accumulate_launch(Nlocal, numVecsProjection, fieldOffset, o_alpha.p, o_xx.p, o_xbar.p, st))) return rc;
    if (!aconj) {
      double* o_rtmp = elliptic->o_z.p;
      if ((rc = accumulate_launch(Nlocal, numVecsProjection, fieldOffset, o_alpha.p, o_bb.p, o_rtmp, st))) return rc;
      return axpby_launch<double>(Nlocal, DevScalar::host(-1.0), o_rtmp, DevScalar::host(1.0), o_r, st);
    }
    if ((rc = matvec(o_bb.p, 0, o_xbar.p, 0))) return rc;
    return axpby_launch<double>(Nlocal, DevScalar::host(-1.0), o_bb.p, DevScalar::host(1.0), o_r, st);
  }
  int post(double* o_x)
  {
    if (timestep < numTimeSteps) return NRSB_OK;
    cudaStream_t st = elliptic->stream;
    int rc;
    const size_t bytes = sizeof(double) * fieldOffset;
    if (numVecsProjection == 0) {
      numVecsProjection = 1;
      NRSB_CUDA(cudaMemcpyAsync(o_xx.p, o_x, bytes, cudaMemcpyDeviceToDevice, st));
    } else if (numVecsProjection == maxNumVecsProjection) {
      numVecsProjection = 1;
      if ((rc = axpby_launch<double>(Nlocal, DevScalar::host(1.0), o_xbar.p, DevScalar::host(1.0), o_x, st))) return rc;
      NRSB_CUDA(cudaMemcpyAsync(o_xx.p, o_x, bytes, cudaMemcpyDeviceToDevice, st));
    } else {
      numVecsProjection++;
      NRSB_CUDA(cudaMemcpyAsync(o_xx.p + (size_t)fieldOffset * (numVecsProjection - 1), o_x, bytes,
                                cudaMemcpyDeviceToDevice, st));
      if ((rc = axpby_launch<double>(Nlocal, DevScalar::host(1.0), o_xbar.p, DevScalar::host(1.0), o_x, st))) return rc;
    }
    const int previous = numVecsProjection;
    const long bOffset = aconj ? 0 : numVecsProjection - 1;
    if ((rc = matvec(o_bb.p, bOffset, o_xx.p, numVecsProjection - 1))) return rc;
    if ((rc = updateProjectionSpace())) return rc;
    if (numVecsProjection < previous) {
      numVecsProjection = 1;
      NRSB_CUDA(cudaMemcpyAsync(o_xx.p, o_x, bytes, cudaMemcpyDeviceToDevice, st));
      if ((rc = matvec(o_bb.p, 0, o_xx.p, 0))) return rc;
      if ((rc = updateProjectionSpace())) return rc;
    }
    return NRSB_OK;
  }
};

}  // namespace nrsb

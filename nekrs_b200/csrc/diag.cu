// diag.cu -- diagonal of the local Helmholtz operator, diag(A_e) = diag(D^T (lambda0 G) D) + lambda1 GwJ.
//
// Replaces kernels/elliptic/ellipticBlockBuildDiagonalHex3D.okl (same argument order: Nelements, Nfields, offset,
// loffset, ggeo, D, lambda0, lambda1, Aq; the reference also passes S = D^T, unused there) as called from
// ellipticUpdateJacobi (ellipticUpdateJacobi.cpp:32-85), and the inversion that follows it
// (linAlg adyMany / padyMany with alpha = 1, linAlg.hpp:127-139).
//
// One block per element, one thread per (i, j) column sweeping k like the reference: G00 of the (j, k) row and G11
// of the (i, k) column come through shared memory, G22 along k and the k-pencil of lambda0 are per-thread registers.
// Same summation order as the OKL thread body (cross terms first, then m = 0..Nq-1 with rr, ss, tt interleaved).
#include "common.cuh"
#include "kernels.hpp"

namespace nrsb {

template <typename T, int Nq, bool kPoisson, bool kLambdaField>
__global__ void __launch_bounds__(Nq* Nq)
    build_diagonal_kernel(const dlong Nelements, const int Nfields, const dlong offset, const dlong loffset,
                          const T* __restrict__ ggeo, const DMat<T, Nq> Dm, const T* __restrict__ lambda0,
                          const T* __restrict__ lambda1, T* __restrict__ Aq)
{
  constexpr int Np = Nq * Nq * Nq;
  __shared__ T s_lambda0[Nq][Nq];
  __shared__ T s_Grr[Nq][Nq];
  __shared__ T s_Gss[Nq][Nq];
  const dlong e = blockIdx.x;
  const int i = threadIdx.x % Nq, j = threadIdx.x / Nq;
  const T* g = ggeo + (size_t)e * 7 * Np;
  for (int l = 0; l < Nfields; ++l) {
    T r_Gtt[Nq], r_lambdat[Nq];
#pragma unroll
    for (int m = 0; m < Nq; ++m) {
      const int n = m * Nq * Nq + j * Nq + i;
      r_Gtt[m] = g[5 * Np + n];
      r_lambdat[m] = lambda0[(kLambdaField ? (size_t)e * Np + n : 0) + (size_t)l * loffset];
    }
#pragma unroll
    for (int k = 0; k < Nq; ++k) {
      const int n = k * Nq * Nq + j * Nq + i;
      __syncthreads();
      s_Grr[j][i] = g[0 * Np + n];
      s_Gss[j][i] = g[2 * Np + n];
      s_lambda0[j][i] = r_lambdat[k];
      __syncthreads();
      const T lbda_0 = r_lambdat[k];
      T r_q = T(0);
      r_q += T(2) * g[1 * Np + n] * lbda_0 * Dm.v[i * Nq + i] * Dm.v[j * Nq + j];
      r_q += T(2) * g[4 * Np + n] * lbda_0 * Dm.v[i * Nq + i] * Dm.v[k * Nq + k];
      r_q += T(2) * g[3 * Np + n] * lbda_0 * Dm.v[j * Nq + j] * Dm.v[k * Nq + k];
#pragma unroll
      for (int m = 0; m < Nq; ++m) {
        r_q += s_Grr[j][m] * s_lambda0[j][m] * Dm.v[m * Nq + i] * Dm.v[m * Nq + i];
        r_q += s_Gss[m][i] * s_lambda0[m][i] * Dm.v[m * Nq + j] * Dm.v[m * Nq + j];
        r_q += r_Gtt[m] * r_lambdat[m] * Dm.v[m * Nq + k] * Dm.v[m * Nq + k];
      }
      if constexpr (!kPoisson) {
        const T lbda_1 = lambda1[(kLambdaField ? (size_t)e * Np + n : 0) + (size_t)l * loffset];
        r_q += g[6 * Np + n] * lbda_1;
      }
      Aq[(size_t)e * Np + n + (size_t)l * offset] = r_q;
    }
  }
  (void)Nelements;
}

template <typename T, int Nq>
static int build_diagonal_nq(dlong Nelements, int Nfields, dlong offset, dlong loffset, const T* ggeo, const T* D_host,
                             const T* lambda0, const T* lambda1, int poisson, int lambdaField, T* Aq,
                             cudaStream_t stream)
{
  DMat<T, Nq> Dm;
  for (int n = 0; n < Nq * Nq; ++n) Dm.v[n] = D_host[n];
#define NRSB_DIAG(P, L)                                                                                        \
  build_diagonal_kernel<T, Nq, P, L><<<Nelements, Nq * Nq, 0, stream>>>(Nelements, Nfields, offset, loffset, ggeo, \
                                                                        Dm, lambda0, lambda1, Aq)
  if (poisson) {
    if (lambdaField)
      NRSB_DIAG(true, true);
    else
      NRSB_DIAG(true, false);
  } else {
    if (lambdaField)
      NRSB_DIAG(false, true);
    else
      NRSB_DIAG(false, false);
  }
#undef NRSB_DIAG
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

template <typename T>
int build_diagonal_launch(int Nq, dlong Nelements, int Nfields, dlong offset, dlong loffset, const T* ggeo,
                          const T* D_host, const T* lambda0, const T* lambda1, int poisson, int lambdaField, T* Aq,
                          cudaStream_t stream)
{
  if (Nelements == 0 || Nfields == 0) return NRSB_OK;
#define NRSB_DIAG_CASE(n)                                                                                         \
  case n:                                                                                                         \
    return build_diagonal_nq<T, n>(Nelements, Nfields, offset, loffset, ggeo, D_host, lambda0, lambda1, poisson,  \
                                   lambdaField, Aq, stream);
  switch (Nq) {
    NRSB_DIAG_CASE(2)
    NRSB_DIAG_CASE(3)
    NRSB_DIAG_CASE(4)
    NRSB_DIAG_CASE(5)
    NRSB_DIAG_CASE(6)
    NRSB_DIAG_CASE(7)
    NRSB_DIAG_CASE(8)
    NRSB_DIAG_CASE(9)
    NRSB_DIAG_CASE(10)
    NRSB_DIAG_CASE(11)
    NRSB_DIAG_CASE(12)
    default:
      set_last_error("ellipticBlockBuildDiagonalHex3D: unsupported Nq (supported: 2..12)");
      return NRSB_ERR_INVALID;
  }
#undef NRSB_DIAG_CASE
}
template int build_diagonal_launch<double>(int, dlong, int, dlong, dlong, const double*, const double*, const double*,
                                           const double*, int, int, double*, cudaStream_t);
template int build_diagonal_launch<float>(int, dlong, int, dlong, dlong, const float*, const float*, const float*,
                                          const float*, int, int, float*, cudaStream_t);

}  // namespace nrsb

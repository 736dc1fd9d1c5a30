// kernels.hpp -- internal launcher prototypes (one per reference kernel family).
#pragma once
#include "common.cuh"

namespace nrsb {

// axhelm.cu
// request for q^T A q from the axhelm launch itself (persistent TMA kernels only): `partials` receives one value per
// CTA, `n` is set to their number; a launcher that cannot honour it sets n = 0 (the caller then uses a separate pass)
struct AxDot {
  double* partials = nullptr;
  int n = 0;
};
template <typename T>
int ax_launch(int Nq, int variant, dlong Nelements, dlong loffset, const dlong* elementList, const T* ggeo,
              const T* D_host, const T* lambda0, const T* lambda1, int poisson, int lambdaField, const T* q, T* Aq,
              cudaStream_t stream, AxDot* dot = nullptr);
template <typename T>
int ax_block_launch(int Nq, int variant, dlong Nelements, int Nfields, dlong offset, dlong loffset,
                    const dlong* elementList, const T* ggeo, const T* D_host, const T* lambda0, const T* lambda1,
                    int lambdaField, const T* q, T* Aq, cudaStream_t stream);
// ellipticStressPartialAxCoeffHex3D (stress.cu): three coupled fields, vgeo = 12 planes per element
template <typename T>
int ax_stress_launch(int Nq, dlong Nelements, dlong offset, dlong loffset, const dlong* elementList, const T* vgeo,
                     const T* D_host, const T* lambda0, const T* lambda1, int lambdaField, const T* q, T* Aq,
                     cudaStream_t stream);
int ax_default_variant(int Nq, int precision);
struct FusedHalo;
template <typename T>
int ax_tma_fused_launch(int Nq, int variant, dlong Nelements, const dlong* elementList, const T* ggeo, const T* D_host,
                        const T* lambda0, const T* lambda1, int poisson, const T* q, T* Aq, const FusedHalo& F,
                        cudaStream_t stream, AxDot* dot = nullptr);

struct FusedRows;
template <typename T>
int ax_tma_gs_launch(int Nq, dlong Nelements, const dlong* elementList, const T* ggeo, const T* D_host,
                     const T* lambda0, const T* lambda1, int poisson, const T* q, T* Aq, const FusedHalo* F,
                     FusedRows* rows, cudaStream_t stream, AxDot* dot = nullptr);

// fdm.cu
int fused_fdm_launch(int Nq, int restrict_, dlong Nelements, const dlong* elementList, float* Su, const float* Sx,
                     const float* Sy, const float* Sz, const float* invL, const float* wts, float* u,
                     cudaStream_t stream);
void set_fdm_variant(int v);
int pre_fdm_launch(int Nq, dlong Nelements, const float* u, float* work1, cudaStream_t stream);
int post_fdm_launch(int Nq, dlong Nelements, const float* work1, const float* work2, float* Su, const float* wts,
                    cudaStream_t stream);

// transfer.cu
int transfer_dispatch(bool coarsen, int NqF, int NqC, dlong Nelements, const float* R_host, const float* in,
                      float* out, cudaStream_t stream);
// diag.cu: ellipticBlockBuildDiagonalHex3D (element diagonal of the Helmholtz operator; D_host row-major [Nq][Nq])
template <typename T>
int build_diagonal_launch(int Nq, dlong Nelements, int Nfields, dlong offset, dlong loffset, const T* ggeo,
                          const T* D_host, const T* lambda0, const T* lambda1, int poisson, int lambdaField, T* Aq,
                          cudaStream_t stream);
bool transfer_supported(int NqF, int NqC);
bool fdm_supported(int Nq);  // extended size Nq + 2 instantiated
int geometric_factors_launch(int Nq, dlong Nelements, const double* d_D, const double* d_gllw, const double* x,
                             const double* y, const double* z, double* ggeo, double* Jac, cudaStream_t stream, double* vgeo = nullptr);

}  // namespace nrsb

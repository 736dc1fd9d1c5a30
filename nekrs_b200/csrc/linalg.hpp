// linalg.hpp -- streaming BLAS-1 and weighted reductions (linAlg_t subset, src/linAlg/linAlg.hpp:33-282).
#pragma once
#include "common.cuh"

namespace nrsb {

// A scalar that is either a host value or lives on the device as scale * num[0] / den[0].
// Krylov coefficients (alpha, beta) are consumed this way so the host never has to read a dot
// product back just to hand it to the next kernel (the reference does a blocking device->host
// copy + MPI_Allreduce for each of the 3 reductions of a PCG iteration, PCG.cpp:55-74).
struct DevScalar {
  double scale = 1.0;
  const double* num = nullptr;
  const double* den = nullptr;
  const double* mul = nullptr;
  double denShift = 0.0;  // value = scale * num * mul / (den + denShift)
  bool guard = false;     // value = 0 when the denominator is not positive (exactly converged CG)
  static DevScalar host(double v)
  {
    DevScalar s;
    s.scale = v;
    return s;
  }
  static DevScalar ratio(const double* n, const double* d, double scale = 1.0, double shift = 0.0)
  {
    DevScalar s;
    s.scale = scale;
    s.num = n;
    s.den = d;
    s.denShift = shift;
    return s;
  }
  __host__ __device__ __forceinline__ double eval() const
  {
    double v = scale;
    if (num) v *= num[0];
    if (mul) v *= mul[0];
    if (den) {
      const double d = den[0] + denShift;
      v = (guard && !(d > 0.0)) ? 0.0 : v / d;
    }
    return v;
  }
};

// cross-rank part of a reduction: one-shot all-reduce through peer-mapped windows (comm.cu)
struct PeerReduce {
  int rank = 0, nranks = 1;
  double* const* slots = nullptr;              // slots[p] -> peer p's window: [parity][nranks][kMaxRed]
  unsigned long long* const* flags = nullptr;  // flags[p] -> peer p's flags:  [nranks]
  unsigned long long* epoch = nullptr;         // local call counter
  int* err = nullptr;                          // host-mapped error word (comm_t::d_err)
};

constexpr int kMaxRed = 16;        // values reduced by one launch
constexpr int kMaxRedBlocks = 1184;  // 148 SMs x 8

// Per-solver reduction workspace (device)
struct ReduceWs {
  double* partials = nullptr;  // [kMaxRedBlocks][kMaxRed]
  unsigned* ticket = nullptr;  // zero between launches
  PeerReduce peer;
};

template <typename T>
int fill_launch(long N, T a, T* x, cudaStream_t s);
template <typename T>
int axpby_launch(long N, DevScalar a, const T* x, DevScalar b, T* y, cudaStream_t s);  // y = a x + b y
template <typename T>
int axpbyz_launch(long N, DevScalar a, const T* x, DevScalar b, const T* y, T* z, cudaStream_t s);
template <typename T>
int axmyz_launch(long N, T a, const T* x, const T* y, T* z, cudaStream_t s);  // z = a x y
template <typename T>
int scale_launch(long N, T a, T* x, cudaStream_t s);
template <typename T>
int add_scalar_launch(long N, DevScalar a, T* x, cudaStream_t s);  // x += a
// "Many" family (linAlg.hpp:75-139): Nfields vectors of N entries, `offset` apart, one launch
template <typename T>
int scale_many_launch(long N, int Nfields, long offset, T a, T* x, cudaStream_t s);
template <typename T>
int axmy_many_launch(long N, int Nfields, long offset, int mode, T a, const T* x, T* y, cudaStream_t s);
template <typename T>
int axmyz_many_launch(long N, int Nfields, long offset, T a, const T* x, const T* y, T* z, cudaStream_t s);
template <typename T>
int ady_many_launch(long N, int Nfields, long offset, T a, T* y, cudaStream_t s);  // y = a / y
template <typename T>
int axdy_launch(long N, T a, const T* x, T* y, cudaStream_t s);  // y = a x / y
template <typename T>
int axpbyz_many_launch(long N, int Nfields, long offset, T a, const T* x, T b, const T* y, T* z, cudaStream_t s);
template <typename T>
int abs_launch(long N, T* x, cudaStream_t s);
int set_scalar_launch(double* dst, DevScalar v, cudaStream_t s);  // dst[0] = v
int copy_d2f_launch(long N, const double* x, float* y, cudaStream_t s);
int copy_f2d_launch(long N, const float* x, double* y, cudaStream_t s);
int axmyz_mixed_launch(long N, float a, const double* x, const float* y, double* z, cudaStream_t s);

// reductions: out[0..nv) on the device (already summed over ranks when ws.peer.nranks > 1)
template <typename T>
int wdot_launch(long N, const T* w, const T* x, const T* y, double* out, const ReduceWs& ws, cudaStream_t s);
template <typename T>
int wnorm2_launch(long N, const T* w, const T* x, double* out, const ReduceWs& ws, cudaStream_t s);
template <typename T>
int sum_launch(long N, const T* x, double* out, const ReduceWs& ws, cudaStream_t s);
int dot_launch(long N, const double* x, const double* y, double* out, const ReduceWs& ws, cudaStream_t s);
// out[v] = sum w x_v y , x_v = X + v*offset, v < NVec <= kMaxRed
int wdot_multi_launch(long N, int NVec, long offset, const double* w, const double* X, const double* y, double* out,
                      const ReduceWs& ws, cudaStream_t s);

// fused PCG update (ellipticBlockUpdatePCG + the axpbyMany that follows it, PCG.cpp:33-83):
//   r -= alpha Ap ; x += alpha p ; out = sum w r^2
int update_pcg_launch(long N, const double* w, const double* Ap, const double* p, DevScalar alpha, double* r,
                      double* x, double* out, const ReduceWs& ws, cudaStream_t s);

// Device-resident PCG control: the same update, but the post step of the reduction turns the sum into the
// residual norm, appends it to a device history, and raises `ctl[0]` once it is <= tol; a raised flag
// freezes x and r in every later launch (iterations launched ahead of the host's convergence check become
// no-ops on the solution).  ctl = {done, iteration at which done was raised, current norm}.
struct PcgControl {
  double* ctl = nullptr;
  double* hist = nullptr;
  double factor = 1.0, tol = 0.0;
  int iter = 0;
  // scalar hand-over to the next iteration, done by the last block instead of two 8-byte device copies:
  // rdotzOld = rdotz ("rdotz2 = rdotz1", PCG.cpp:121) and, without preconditioner, rdotz = the new residual
  // norm ("rdotz1 = rdotr", PCG.cpp:131)
  double* rdotz = nullptr;
  double* rdotzOld = nullptr;
  int normIsRdotz = 0;
};
int update_pcg_ctl_launch(long N, const double* w, const double* Ap, const double* p, DevScalar alpha, double* r,
                          double* x, double* out, PcgControl c, const ReduceWs& ws, cudaStream_t s);
// out = sum w x y ; alpha[0] = num[0] / (out + shift)   (pAp and the step length in one launch)
int sum_ratio_launch(long n, const double* v, double* out, const double* num, double shift, double* alpha,
                     const ReduceWs& ws, cudaStream_t s);
int wdot_ratio_launch(long N, const double* w, const double* x, const double* y, double* out, const double* num,
                      double shift, double* alpha, const ReduceWs& ws, cudaStream_t s);

// Chebyshev updates (updateChebyshev.okl, updateFourthKindChebyshev.okl)
int update_chebyshev_launch(long N, float dCoeff, float rCoeff, const float* SAd, float* d, float* r, float* x,
                            cudaStream_t s);
int update_fourth_chebyshev_launch(long N, float beta, const float* Ad, const float* d, float* r, float* x,
                                   cudaStream_t s);

// GMRES (gramSchmidtOrthogonalization.c, updatePGMRESSolution.c, fusedResidualAndNorm.c)
int gram_schmidt_launch(long N, long offset, int gmresSize, const double* w, const double* y, const double* V,
                        double* wv, double* out, const ReduceWs& ws, cudaStream_t s);
int update_pgmres_solution_launch(long N, long offset, int gmresSize, const double* y, const double* Z, double* x,
                                  cudaStream_t s);
int fused_residual_and_norm_launch(long N, const double* w, const double* b, const double* Ax, double* r, double* out,
                                   const ReduceWs& ws, cudaStream_t s);

}  // namespace nrsb

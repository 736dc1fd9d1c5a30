// linalg.cu -- streaming BLAS-1 kernels and single-pass weighted reductions.
//
// Replaces: kernels/linAlg/*.okl (fill, axpby(Many), axmyz, axmy, scale(Many), add, sum,
// weightedInnerProd(Many), weightedNorm2(Many), weightedInnerProdMulti, innerProd) and
// kernels/elliptic/{ellipticBlockUpdatePCG, updateChebyshev, updateFourthKindChebyshev,
// gramSchmidtOrthogonalization, updatePGMRESSolution, fusedResidualAndNorm}.okl, plus
// kernels/core/copy{Dfloat,Pfloat}To{Pfloat,Dfloat}.okl.
//
// Reductions: the reference writes one partial per 256-thread block, copies all partials to the
// host, sums them there and calls MPI_Allreduce (linAlg.cpp:986-1034, PCG.cpp:55-74).  Here one
// launch does everything: grid-stride accumulation in fp64, warp-shuffle + shared-memory block
// reduction, per-block partials, and the LAST block to finish (atomic ticket) folds the partials
// in a fixed order, performs the cross-GPU one-shot all-reduce through peer-mapped windows
// (NVLink stores + flag, comm.cu) and leaves the scalar on the device.  The result is
// deterministic for a given (N, grid) and bit-identical on every rank.
#include "linalg.hpp"
#include "reduce.cuh"

namespace nrsb {

static inline int stream_grid(long N, int perThread = 4)
{
  long b = (N + (long)kBlockSize * perThread - 1) / ((long)kBlockSize * perThread);
  const long cap = (long)kNumSMs * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ----------------------------------------------------------------------------- elementwise
template <typename F>
__global__ void __launch_bounds__(kBlockSize) ew_kernel(long N, F f)
{
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) f(i);
}

template <typename F>
static int ew_launch(long N, F f, cudaStream_t s)
{
  if (N <= 0) return NRSB_OK;
  ew_kernel<<<stream_grid(N), kBlockSize, 0, s>>>(N, f);
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

template <typename T>
int fill_launch(long N, T a, T* x, cudaStream_t s)
{
  return ew_launch(N, [=] __device__(long i) { x[i] = a; }, s);
}
template <typename T>
int axpby_launch(long N, DevScalar a, const T* x, DevScalar b, T* y, cudaStream_t s)
{
  const bool bZero = (b.num == nullptr && b.den == nullptr && b.scale == 0.0);
  if (bZero)  // y may be uninitialised (linAlg axpby with beta = 0 is used as a scaled copy)
    return ew_launch(N, [=] __device__(long i) { const T av = (T)a.eval(); y[i] = av * x[i]; }, s);
  return ew_launch(
      N,
      [=] __device__(long i) {
        const T av = (T)a.eval(), bv = (T)b.eval();
        y[i] = av * x[i] + bv * y[i];
      },
      s);
}
template <typename T>
int axpbyz_launch(long N, DevScalar a, const T* x, DevScalar b, const T* y, T* z, cudaStream_t s)
{
  return ew_launch(
      N,
      [=] __device__(long i) {
        const T av = (T)a.eval(), bv = (T)b.eval();
        z[i] = av * x[i] + bv * y[i];
      },
      s);
}
template <typename T>
int axmyz_launch(long N, T a, const T* x, const T* y, T* z, cudaStream_t s)
{
  return ew_launch(N, [=] __device__(long i) { z[i] = a * x[i] * y[i]; }, s);
}
template <typename T>
int scale_launch(long N, T a, T* x, cudaStream_t s)
{
  return ew_launch(N, [=] __device__(long i) { x[i] *= a; }, s);
}
// ---- the "Many" family (linAlg.hpp:75-139): Nfields vectors of N entries, `offset` apart, in ONE launch.
//      The index runs over N * Nfields; field-major so that consecutive threads touch consecutive addresses.
template <typename F>
static int ew_many_launch(long N, int Nfields, F f, cudaStream_t s)
{
  if (N <= 0 || Nfields <= 0) return NRSB_OK;
  return ew_launch(N * Nfields, [=] __device__(long g) { f(g % N, (int)(g / N)); }, s);
}
template <typename T>
int scale_many_launch(long N, int Nfields, long offset, T a, T* x, cudaStream_t s)
{
  return ew_many_launch(N, Nfields, [=] __device__(long i, int fld) { x[i + fld * offset] *= a; }, s);
}
// mode 1: y[n,fld] = a x[n,fld] y[n,fld] ; mode 0: y[n,fld] = a x[n] y[n,fld]   (linAlg.hpp:106-113)
template <typename T>
int axmy_many_launch(long N, int Nfields, long offset, int mode, T a, const T* x, T* y, cudaStream_t s)
{
  return ew_many_launch(
      N, Nfields,
      [=] __device__(long i, int fld) { y[i + fld * offset] = a * x[i + (mode ? fld * offset : 0)] * y[i + fld * offset]; },
      s);
}
template <typename T>
int axmyz_many_launch(long N, int Nfields, long offset, T a, const T* x, const T* y, T* z, cudaStream_t s)
{
  return ew_many_launch(
      N, Nfields, [=] __device__(long i, int fld) { z[i + fld * offset] = a * x[i + fld * offset] * y[i + fld * offset]; },
      s);
}
// y = a / y  (ady / adyMany / padyMany, linAlg.hpp:133-139)
template <typename T>
int ady_many_launch(long N, int Nfields, long offset, T a, T* y, cudaStream_t s)
{
  return ew_many_launch(N, Nfields, [=] __device__(long i, int fld) { y[i + fld * offset] = a / y[i + fld * offset]; }, s);
}
// y = a x / y  (axdy, linAlg.hpp:141-142)
template <typename T>
int axdy_launch(long N, T a, const T* x, T* y, cudaStream_t s)
{
  return ew_launch(N, [=] __device__(long i) { y[i] = a * x[i] / y[i]; }, s);
}
template <typename T>
int axpbyz_many_launch(long N, int Nfields, long offset, T a, const T* x, T b, const T* y, T* z, cudaStream_t s)
{
  return ew_many_launch(
      N, Nfields,
      [=] __device__(long i, int fld) { z[i + fld * offset] = a * x[i + fld * offset] + b * y[i + fld * offset]; }, s);
}
template <typename T>
int abs_launch(long N, T* x, cudaStream_t s)
{
  return ew_launch(N, [=] __device__(long i) { x[i] = x[i] < T(0) ? -x[i] : x[i]; }, s);
}

template <typename T>
int add_scalar_launch(long N, DevScalar a, T* x, cudaStream_t s)
{
  return ew_launch(N, [=] __device__(long i) { x[i] += (T)a.eval(); }, s);
}
int set_scalar_launch(double* dst, DevScalar v, cudaStream_t s)
{
  return ew_launch(1, [=] __device__(long i) { dst[i] = v.eval(); }, s);
}
int copy_d2f_launch(long N, const double* x, float* y, cudaStream_t s)
{
  return ew_launch(N, [=] __device__(long i) { y[i] = (float)x[i]; }, s);
}
int copy_f2d_launch(long N, const float* x, double* y, cudaStream_t s)
{
  return ew_launch(N, [=] __device__(long i) { y[i] = (double)x[i]; }, s);
}
int axmyz_mixed_launch(long N, float a, const double* x, const float* y, double* z, cudaStream_t s)
{
  // axmyzManyPfloat.c: z = dfloat(alpha * x * y)
  return ew_launch(N, [=] __device__(long i) { z[i] = (double)(a * x[i] * y[i]); }, s);
}
int update_chebyshev_launch(long N, float dCoeff, float rCoeff, const float* SAd, float* d, float* r, float* x,
                            cudaStream_t s)
{
  return ew_launch(
      N,
      [=] __device__(long i) {
        const float dn = d[i];
        const float rn = r[i] - SAd[i];
        x[i] = x[i] + dn;
        r[i] = rn;
        d[i] = dCoeff * dn + rCoeff * rn;
      },
      s);
}
int update_fourth_chebyshev_launch(long N, float beta, const float* Ad, const float* d, float* r, float* x,
                                   cudaStream_t s)
{
  return ew_launch(
      N,
      [=] __device__(long i) {
        x[i] = x[i] + beta * d[i];
        r[i] = r[i] - Ad[i];
      },
      s);
}
int update_pgmres_solution_launch(long N, long offset, int gmresSize, const double* y, const double* Z, double* x,
                                  cudaStream_t s)
{
  return ew_launch(
      N,
      [=] __device__(long i) {
        double xv = x[i];
        for (int j = 0; j < gmresSize; ++j) xv += Z[i + (size_t)j * offset] * y[j];
        x[i] = xv;
      },
      s);
}

// accumulate.okl / multiScaledAddwOffset.okl (solution projection), Nfields = 1
int accumulate_launch(long N, int m, long fieldOffset, const double* alpha, const double* x, double* y,
                      cudaStream_t s)
{
  return ew_launch(
      N,
      [=] __device__(long n) {
        double v = alpha[0] * x[n];
        for (int k = 1; k < m; ++k) v += alpha[k] * x[n + (size_t)k * fieldOffset];
        y[n] = v;
      },
      s);
}
int multi_scaled_add_w_offset_launch(long N, int m, long destOffset, long fieldOffset, const double* alphas,
                                     double beta, double* x, cudaStream_t s)
{
  return ew_launch(
      N,
      [=] __device__(long n) {
        double v = x[n + destOffset];
        for (int k = 0; k < m - 1; ++k) v = -alphas[k] * x[n + (size_t)k * fieldOffset] + beta * v;
        x[n + destOffset] = v;
      },
      s);
}

#define NRSB_INST(T)                                                                                  \
  template int fill_launch<T>(long, T, T*, cudaStream_t);                                             \
  template int axpby_launch<T>(long, DevScalar, const T*, DevScalar, T*, cudaStream_t);               \
  template int axpbyz_launch<T>(long, DevScalar, const T*, DevScalar, const T*, T*, cudaStream_t);    \
  template int axmyz_launch<T>(long, T, const T*, const T*, T*, cudaStream_t);                        \
  template int scale_launch<T>(long, T, T*, cudaStream_t);                                            \
  template int add_scalar_launch<T>(long, DevScalar, T*, cudaStream_t);                               \
  template int scale_many_launch<T>(long, int, long, T, T*, cudaStream_t);                            \
  template int axmy_many_launch<T>(long, int, long, int, T, const T*, T*, cudaStream_t);              \
  template int axmyz_many_launch<T>(long, int, long, T, const T*, const T*, T*, cudaStream_t);        \
  template int ady_many_launch<T>(long, int, long, T, T*, cudaStream_t);                              \
  template int axdy_launch<T>(long, T, const T*, T*, cudaStream_t);                                   \
  template int axpbyz_many_launch<T>(long, int, long, T, const T*, T, const T*, T*, cudaStream_t);    \
  template int abs_launch<T>(long, T*, cudaStream_t);
NRSB_INST(double)
NRSB_INST(float)
#undef NRSB_INST

// ----------------------------------------------------------------------------- reductions
// (template in reduce.cuh)

template <typename T>
int wdot_launch(long N, const T* w, const T* x, const T* y, double* out, const ReduceWs& ws, cudaStream_t s)
{
  return reduce_launch<1>(
      N, [=] __device__(long i, double* acc) { acc[0] += (double)x[i] * (double)y[i] * (double)w[i]; }, 1, out, ws,
      s);
}
template <typename T>
int wnorm2_launch(long N, const T* w, const T* x, double* out, const ReduceWs& ws, cudaStream_t s)
{
  return reduce_launch<1>(
      N,
      [=] __device__(long i, double* acc) {
        const double xv = (double)x[i];
        acc[0] += xv * xv * (double)w[i];
      },
      1, out, ws, s);
}
template <typename T>
int sum_launch(long N, const T* x, double* out, const ReduceWs& ws, cudaStream_t s)
{
  return reduce_launch<1>(N, [=] __device__(long i, double* acc) { acc[0] += (double)x[i]; }, 1, out, ws, s);
}
int dot_launch(long N, const double* x, const double* y, double* out, const ReduceWs& ws, cudaStream_t s)
{
  return reduce_launch<1>(N, [=] __device__(long i, double* acc) { acc[0] += x[i] * y[i]; }, 1, out, ws, s);
}
template int wdot_launch<double>(long, const double*, const double*, const double*, double*, const ReduceWs&,
                                 cudaStream_t);
template int wdot_launch<float>(long, const float*, const float*, const float*, double*, const ReduceWs&,
                                cudaStream_t);
template int wnorm2_launch<double>(long, const double*, const double*, double*, const ReduceWs&, cudaStream_t);
template int wnorm2_launch<float>(long, const float*, const float*, double*, const ReduceWs&, cudaStream_t);
template int sum_launch<double>(long, const double*, double*, const ReduceWs&, cudaStream_t);
template int sum_launch<float>(long, const float*, double*, const ReduceWs&, cudaStream_t);

int wdot_multi_launch(long N, int NVec, long offset, const double* w, const double* X, const double* y, double* out,
                      const ReduceWs& ws, cudaStream_t s)
{
  if (NVec < 1 || NVec > kMaxRed) {
    set_last_error("weightedInnerProdMulti: NVec must be in 1..16");
    return NRSB_ERR_INVALID;
  }
  // two nodes per step with 16-byte loads: at 85 registers only 512 threads are resident per SM, and one 8-byte
  // load per vector and thread does not keep enough bytes in flight (measured in the BPS5 launch list: 2.5 TB/s
  // against 4.6 TB/s for the Gram-Schmidt update over the same vectors)
  const bool pairs = N % 2 == 0 && offset % 2 == 0 && (((uintptr_t)w | (uintptr_t)X | (uintptr_t)y) & 15) == 0;
  if (pairs) {
    const double2* w2 = reinterpret_cast<const double2*>(w);
    const double2* y2 = reinterpret_cast<const double2*>(y);
    const double2* X2 = reinterpret_cast<const double2*>(X);
    const long off2 = offset / 2;
    auto op2 = [=] __device__(long i, double* acc) {
      const double2 a = w2[i], b = y2[i];
      const double wy0 = a.x * b.x, wy1 = a.y * b.y;
#pragma unroll
      for (int v = 0; v < kMaxRed; ++v)
        if (v < NVec) {
          const double2 x = X2[i + (size_t)v * off2];
          acc[v] += wy0 * x.x;  // ascending node order within the thread, as the one-node form
          acc[v] += wy1 * x.y;
        }
    };
    return reduce_launch<kMaxRed>(N / 2, op2, NVec, out, ws, s);
  }
  auto op = [=] __device__(long i, double* acc) {
    const double wy = w[i] * y[i];
#pragma unroll
    for (int v = 0; v < kMaxRed; ++v)
      if (v < NVec) acc[v] += wy * X[i + (size_t)v * offset];
  };
  return reduce_launch<kMaxRed>(N, op, NVec, out, ws, s);
}

int update_pcg_launch(long N, const double* w, const double* Ap, const double* p, DevScalar alpha, double* r,
                      double* x, double* out, const ReduceWs& ws, cudaStream_t s)
{
  auto op = [=] __device__(long i, double* acc) {
    const double a = alpha.eval();
    const double rn = r[i] - a * Ap[i];
    r[i] = rn;
    if (x) x[i] = a * p[i] + x[i];
    acc[0] += rn * rn * w[i];
  };
  return reduce_launch<1>(N, op, 1, out, ws, s);
}

namespace {
struct PcgCtlPost {
  PcgControl c;
  __device__ __forceinline__ void operator()(double* tot) const
  {
    if (c.ctl[0] != 0.0) return;  // frozen
    const double rd = sqrt(tot[0] * c.factor);
    c.ctl[2] = rd;
    c.hist[c.iter - 1] = rd;
    if (c.rdotzOld) c.rdotzOld[0] = c.rdotz[0];
    if (c.normIsRdotz) c.rdotz[0] = rd;
    if (!(rd > c.tol)) {  // also taken for NaN, like the reference's loop condition
      c.ctl[0] = 1.0;
      c.ctl[1] = (double)c.iter;
    }
  }
};
struct RatioPost {
  const double* num;
  double shift;
  double* alpha;
  __device__ __forceinline__ void operator()(double* tot) const { alpha[0] = num[0] / (tot[0] + shift); }
};
}  // namespace

int update_pcg_ctl_launch(long N, const double* w, const double* Ap, const double* p, DevScalar alpha, double* r,
                          double* x, double* out, PcgControl c, const ReduceWs& ws, cudaStream_t s)
{
  const double* done = c.ctl;
  auto op = [=] __device__(long i, double* acc) {
    if (__ldg(done) != 0.0) return;
    const double a = alpha.eval();
    const double rn = r[i] - a * Ap[i];
    r[i] = rn;
    x[i] = a * p[i] + x[i];
    acc[0] += rn * rn * w[i];
  };
  return reduce_launch<1>(N, op, 1, out, ws, s, PcgCtlPost{c});
}

int wdot_ratio_launch(long N, const double* w, const double* x, const double* y, double* out, const double* num,
                      double shift, double* alpha, const ReduceWs& ws, cudaStream_t s)
{
  return reduce_launch<1>(
      N, [=] __device__(long i, double* acc) { acc[0] += x[i] * y[i] * w[i]; }, 1, out, ws, s,
      RatioPost{num, shift, alpha});
}

// out = sum_i v[i] (fixed order), alpha = num / (out + shift): folds per-CTA partials left by another kernel
int sum_ratio_launch(long n, const double* v, double* out, const double* num, double shift, double* alpha,
                     const ReduceWs& ws, cudaStream_t s)
{
  return reduce_launch<1>(
      n, [=] __device__(long i, double* acc) { acc[0] += v[i]; }, 1, out, ws, s, RatioPost{num, shift, alpha});
}

int gram_schmidt_launch(long N, long offset, int gmresSize, const double* w, const double* y, const double* V,
                        double* wv, double* out, const ReduceWs& ws, cudaStream_t s)
{
  auto op = [=] __device__(long i, double* acc) {
    double v = wv[i];
    for (int j = 0; j < gmresSize; ++j) v -= y[j] * V[i + (size_t)j * offset];
    wv[i] = v;
    acc[0] += v * v * w[i];
  };
  return reduce_launch<1>(N, op, 1, out, ws, s);
}

int fused_residual_and_norm_launch(long N, const double* w, const double* b, const double* Ax, double* r, double* out,
                                   const ReduceWs& ws, cudaStream_t s)
{
  auto op = [=] __device__(long i, double* acc) {
    const double rn = b[i] - Ax[i];
    r[i] = rn;
    acc[0] += rn * rn * w[i];
  };
  return reduce_launch<1>(N, op, 1, out, ws, s);
}

}  // namespace nrsb

// transfer.cu -- p-multigrid restriction / prolongation and geometric factors.
//
// Replaces kernels/elliptic/ellipticPreconCoarsenHex3D.okl:26-105 (serial .c:26-100),
// ellipticPreconProlongateHex3D.okl (serial .c:26-110) and, for setup,
// kernels/mesh/geometricFactorsHex3D.okl:26-142.
//
//   coarsen   : qc  = (R x R x R) qf          R[NqC][NqF]  (= interpolation^T)
//   prolongate: qN += (R^T x R^T x R^T) qc
// pfloat in/out, dfloat accumulation (the reference keeps dfloat=double inside these kernels,
// registerEllipticPreconditionerKernels.cpp:252-276).
//
// R is passed by value (constant bank); the three 1-D contractions go through shared memory in
// fp64.  One element per NqF x NqF thread slab.
#include "common.cuh"

namespace nrsb {

template <int NqF, int NqC>
struct RMat {
  float v[NqC * NqF];
};

template <int NqF, int NqC>
__global__ void __launch_bounds__(NqF* NqF)
    coarsen_kernel(const dlong Nelements, const RMat<NqF, NqC> R, const float* __restrict__ qf, float* __restrict__ qc)
{
  constexpr int NpF = NqF * NqF * NqF, NpC = NqC * NqC * NqC;
  __shared__ double s_a[NqC][NqF][NqF];  // after z contraction: [kc][j][i]
  __shared__ double s_b[NqC][NqC][NqF];  // after y contraction: [kc][jc][i]
  const dlong e = blockIdx.x;
  const int t = threadIdx.x;
  const int i = t % NqF, j = t / NqF;
  // z: thread (i,j) owns the k pencil (coalesced global reads)
  {
    double acc[NqC];
#pragma unroll
    for (int m = 0; m < NqC; ++m) acc[m] = 0;
#pragma unroll
    for (int k = 0; k < NqF; ++k) {
      const double v = qf[(size_t)e * NpF + k * NqF * NqF + t];
#pragma unroll
      for (int m = 0; m < NqC; ++m) acc[m] += (double)R.v[m * NqF + k] * v;
    }
#pragma unroll
    for (int m = 0; m < NqC; ++m) s_a[m][j][i] = acc[m];
  }
  __syncthreads();
  // y: threads (i, kc) own the j pencil
  for (int p = t; p < NqF * NqC; p += NqF * NqF) {
    const int ii = p % NqF, kc = p / NqF;
    double v[NqF];
#pragma unroll
    for (int m = 0; m < NqF; ++m) v[m] = s_a[kc][m][ii];
#pragma unroll
    for (int jc = 0; jc < NqC; ++jc) {
      double acc = 0;
#pragma unroll
      for (int m = 0; m < NqF; ++m) acc += (double)R.v[jc * NqF + m] * v[m];
      s_b[kc][jc][ii] = acc;
    }
  }
  __syncthreads();
  // x: threads (jc, kc) own the i pencil
  for (int p = t; p < NqC * NqC; p += NqF * NqF) {
    const int jc = p % NqC, kc = p / NqC;
    double v[NqF];
#pragma unroll
    for (int m = 0; m < NqF; ++m) v[m] = s_b[kc][jc][m];
#pragma unroll
    for (int ic = 0; ic < NqC; ++ic) {
      double acc = 0;
#pragma unroll
      for (int m = 0; m < NqF; ++m) acc += (double)R.v[ic * NqF + m] * v[m];
      qc[(size_t)e * NpC + kc * NqC * NqC + jc * NqC + ic] = (float)acc;
    }
  }
}

template <int NqF, int NqC>
__global__ void __launch_bounds__(NqF* NqF)
    prolongate_kernel(const dlong Nelements, const RMat<NqF, NqC> R, const float* __restrict__ qc,
                      float* __restrict__ qN)
{
  constexpr int NpF = NqF * NqF * NqF, NpC = NqC * NqC * NqC;
  __shared__ double s_c[NqC][NqC][NqC];
  __shared__ double s_a[NqC][NqC][NqF];  // after x: [kc][jc][i]
  __shared__ double s_b[NqC][NqF][NqF];  // after y: [kc][j][i]
  const dlong e = blockIdx.x;
  const int t = threadIdx.x;
  const int i = t % NqF, j = t / NqF;
  for (int p = t; p < NpC; p += NqF * NqF) (&s_c[0][0][0])[p] = qc[(size_t)e * NpC + p];
  __syncthreads();
  // x: threads (jc,kc) own the ic pencil
  for (int p = t; p < NqC * NqC; p += NqF * NqF) {
    const int jc = p % NqC, kc = p / NqC;
    double v[NqC];
#pragma unroll
    for (int m = 0; m < NqC; ++m) v[m] = s_c[kc][jc][m];
#pragma unroll
    for (int ii = 0; ii < NqF; ++ii) {
      double acc = 0;
#pragma unroll
      for (int m = 0; m < NqC; ++m) acc += (double)R.v[m * NqF + ii] * v[m];
      s_a[kc][jc][ii] = acc;
    }
  }
  __syncthreads();
  // y: threads (i,kc) own the jc pencil
  for (int p = t; p < NqF * NqC; p += NqF * NqF) {
    const int ii = p % NqF, kc = p / NqF;
    double v[NqC];
#pragma unroll
    for (int m = 0; m < NqC; ++m) v[m] = s_a[kc][m][ii];
#pragma unroll
    for (int jj = 0; jj < NqF; ++jj) {
      double acc = 0;
#pragma unroll
      for (int m = 0; m < NqC; ++m) acc += (double)R.v[m * NqF + jj] * v[m];
      s_b[kc][jj][ii] = acc;
    }
  }
  __syncthreads();
  // z: thread (i,j) owns the kc pencil; coalesced read-modify-write of the fine vector
  {
    double v[NqC];
#pragma unroll
    for (int m = 0; m < NqC; ++m) v[m] = s_b[m][j][i];
#pragma unroll
    for (int k = 0; k < NqF; ++k) {
      double acc = 0;
#pragma unroll
      for (int m = 0; m < NqC; ++m) acc += (double)R.v[m * NqF + k] * v[m];
      float* dst = qN + (size_t)e * NpF + k * NqF * NqF + t;
      *dst = (float)((double)*dst + acc);
    }
  }
}

template <int NqF, int NqC>
static int transfer_launch(bool coarsen, dlong Nelements, const float* R_host, const float* in, float* out,
                           cudaStream_t stream)
{
  RMat<NqF, NqC> R;
  for (int n = 0; n < NqF * NqC; ++n) R.v[n] = R_host[n];
  if (coarsen)
    coarsen_kernel<NqF, NqC><<<Nelements, NqF * NqF, 0, stream>>>(Nelements, R, in, out);
  else
    prolongate_kernel<NqF, NqC><<<Nelements, NqF * NqF, 0, stream>>>(Nelements, R, in, out);
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

// is the (NqFine, NqCoarse) pair instantiated?  (setup-time check + CPU test against determineMGLevels)
bool transfer_supported(int NqF, int NqC)
{
  static const int pairs[][2] = {{3, 2},  {4, 2},  {4, 3},  {5, 2},  {5, 3},  {5, 4},  {6, 2},   {6, 3},  {6, 4},
                                 {6, 5},  {7, 2},  {7, 4},  {7, 5},  {7, 6},  {8, 2},  {8, 4},   {8, 5},  {8, 6},
                                 {8, 7},  {9, 2},  {9, 4},  {9, 6},  {9, 7},  {9, 8},  {10, 2},  {10, 4}, {10, 6},
                                 {10, 8}, {10, 9}, {11, 2}, {11, 6}, {11, 7}, {11, 8}, {11, 9},  {11, 10}, {12, 2},
                                 {12, 6}, {12, 7}, {12, 8}, {12, 10}, {12, 11}};
  for (auto& p : pairs)
    if (p[0] == NqF && p[1] == NqC) return true;
  return false;
}

int transfer_dispatch(bool coarsen, int NqF, int NqC, dlong Nelements, const float* R_host, const float* in,
                      float* out, cudaStream_t stream)
{
  if (Nelements == 0) return NRSB_OK;
#define TR(f, c) \
  if (NqF == f && NqC == c) return transfer_launch<f, c>(coarsen, Nelements, R_host, in, out, stream);
  // level pairs of determineMGLevels.cpp:58-95 (N -> next coarser N) plus user schedules seen in examples
  TR(3, 2) TR(4, 2) TR(4, 3) TR(5, 2) TR(5, 3) TR(5, 4) TR(6, 2) TR(6, 3) TR(6, 4) TR(6, 5)
  TR(7, 2) TR(7, 4) TR(7, 5) TR(7, 6) TR(8, 2) TR(8, 4) TR(8, 5) TR(8, 6) TR(8, 7)
  TR(9, 2) TR(9, 4) TR(9, 6) TR(9, 7) TR(9, 8) TR(10, 2) TR(10, 4) TR(10, 6) TR(10, 8) TR(10, 9)
  TR(11, 2) TR(11, 6) TR(11, 7) TR(11, 8) TR(11, 9) TR(11, 10)
  TR(12, 2) TR(12, 6) TR(12, 7) TR(12, 8) TR(12, 10) TR(12, 11)
#undef TR
  set_last_error("coarsen/prolongate: unsupported (NqFine, NqCoarse) pair");
  return NRSB_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------
// geometricFactorsHex3D (setup): one thread per node, run-time Nq
__global__ void __launch_bounds__(256)
    geometric_factors_kernel(const dlong Nelements, const int Nq, const double* __restrict__ D,
                             const double* __restrict__ gllw, const double* __restrict__ x,
                             const double* __restrict__ y, const double* __restrict__ z, double* __restrict__ ggeo,
                             double* __restrict__ Jac, double* __restrict__ vgeo)
{
  const int Np = Nq * Nq * Nq;
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long)Nelements * Np) return;
  const dlong e = gid / Np;
  const int n = gid % Np;
  const int i = n % Nq, j = (n / Nq) % Nq, k = n / (Nq * Nq);
  const double *xe = x + (size_t)e * Np, *ye = y + (size_t)e * Np, *ze = z + (size_t)e * Np;
  double xr = 0, yr = 0, zr = 0, xs = 0, ys = 0, zs = 0, xt = 0, yt = 0, zt = 0;
  for (int m = 0; m < Nq; ++m) {
    const double Dim = D[i * Nq + m], Djm = D[j * Nq + m], Dkm = D[k * Nq + m];
    const int r = k * Nq * Nq + j * Nq + m, s = k * Nq * Nq + m * Nq + i, t = m * Nq * Nq + j * Nq + i;
    xr += Dim * xe[r];
    xs += Djm * xe[s];
    xt += Dkm * xe[t];
    yr += Dim * ye[r];
    ys += Djm * ye[s];
    yt += Dkm * ye[t];
    zr += Dim * ze[r];
    zs += Djm * ze[s];
    zt += Dkm * ze[t];
  }
  const double J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt);
  const double Jinv = 1. / J;
  const double JW = J * gllw[i] * gllw[j] * gllw[k];
  const double rx = (ys * zt - zs * yt) * Jinv, ry = -(xs * zt - zs * xt) * Jinv, rz = (xs * yt - ys * xt) * Jinv;
  const double sx = -(yr * zt - zr * yt) * Jinv, sy = (xr * zt - zr * xt) * Jinv, sz = -(xr * yt - yr * xt) * Jinv;
  const double tx = (yr * zs - zr * ys) * Jinv, ty = -(xr * zs - zr * xs) * Jinv, tz = (xr * ys - yr * xs) * Jinv;
  if (Jac) Jac[(size_t)e * Np + n] = J;
  if (vgeo) {  // mesh->vgeo: rx,ry,rz,sx,sy,sz,tx,ty,tz,J,JW,1/JW (mesh3D.h:82-93, meshGeometricFactorsHex3D.cpp)
    double* v = vgeo + (size_t)12 * Np * e + n;
    const double f[12] = {rx, ry, rz, sx, sy, sz, tx, ty, tz, J, JW, 1. / JW};
    for (int c = 0; c < 12; ++c) v[c * (size_t)Np] = f[c];
  }
  if (!ggeo) return;
  double* g = ggeo + (size_t)7 * Np * e + n;
  g[0 * (size_t)Np] = JW * (rx * rx + ry * ry + rz * rz);
  g[1 * (size_t)Np] = JW * (rx * sx + ry * sy + rz * sz);
  g[4 * (size_t)Np] = JW * (rx * tx + ry * ty + rz * tz);
  g[2 * (size_t)Np] = JW * (sx * sx + sy * sy + sz * sz);
  g[3 * (size_t)Np] = JW * (sx * tx + sy * ty + sz * tz);
  g[5 * (size_t)Np] = JW * (tx * tx + ty * ty + tz * tz);
  g[6 * (size_t)Np] = JW;
}

int geometric_factors_launch(int Nq, dlong Nelements, const double* d_D, const double* d_gllw, const double* x,
                             const double* y, const double* z, double* ggeo, double* Jac, cudaStream_t stream,
                             double* vgeo)
{
  const long total = (long)Nelements * Nq * Nq * Nq;
  if (total == 0) return NRSB_OK;
  geometric_factors_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(Nelements, Nq, d_D, d_gllw, x, y, z,
                                                                                ggeo, Jac, vgeo);
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

}  // namespace nrsb

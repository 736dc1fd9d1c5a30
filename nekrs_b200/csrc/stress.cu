// stress.cu -- the coupled three-field viscous-stress operator
//     A (u,v,w) = -div( lambda0 (grad q + grad q^T) ) + lambda1 q      (weak form, collocated JW)
// ellipticStressPartialAxCoeffHex3D (kernels/elliptic/ellipticStressPartialAxCoeffHex3D.okl, serial twin .c:1-169),
// the AxKernel of a block solver with stressForm (ellipticSetup.cpp:240-249).
//
// vgeo: 12 planes per element, rx,ry,rz,sx,sy,sz,tx,ty,tz = 0..8, J = 9, JW = 10, 1/JW = 11 (mesh3D.h:82-93); ten
// of them are read.  Algorithmic bytes per element: (10 + 3 + 3) Np w = 65 536 B at N = 7 in fp64.
//
// One element per block of Nq x Nq threads sweeping k.  The three fields and, after the first sweep, their nine
// reference-space fluxes live in shared memory (12 Np values: 48 KB at N = 7 in fp64); the geometric factors and
// coefficients of a node are read once, straight into registers, by the thread that owns the node in both sweeps.
#include "common.cuh"
#include "kernels.hpp"

namespace nrsb {

namespace {

template <typename T, int Nq, bool kLambdaField>
__global__ void __launch_bounds__(Nq* Nq)
    ax_stress_kernel(const dlong Nelements, const dlong offset, const dlong loffset,
                     const dlong* __restrict__ elementList, const T* __restrict__ vgeo, const DMat<T, Nq> Dm,
                     const T* __restrict__ lambda0, const T* __restrict__ lambda1, const T* __restrict__ q,
                     T* __restrict__ Aq)
{
  constexpr int Nq2 = Nq * Nq;
  constexpr int Np = Nq2 * Nq;
  extern __shared__ __align__(16) unsigned char smraw[];
  T* s_q = reinterpret_cast<T*>(smraw);  // [3][Np]
  T* s_F = s_q + 3 * Np;                 // [3 fields][r,s,t][Np]
  __shared__ T s_D[Nq2];

  const dlong e = elementList[blockIdx.x];
  const int t = threadIdx.x;
  const int i = t % Nq, j = t / Nq;
  s_D[t] = Dm.v[t];
#pragma unroll
  for (int f = 0; f < 3; ++f)
#pragma unroll
    for (int k = 0; k < Nq; ++k) s_q[f * Np + k * Nq2 + t] = q[(size_t)e * Np + k * Nq2 + t + (size_t)f * offset];
  __syncthreads();

  const T* g = vgeo + (size_t)e * Np * 12;
  T lam0c[3] = {T(0), T(0), T(0)};
  if (!kLambdaField)
#pragma unroll
    for (int f = 0; f < 3; ++f) lam0c[f] = lambda0[(size_t)f * loffset];

  // sweep 1: derivatives -> physical gradients -> stress -> reference-space fluxes
#pragma unroll 1
  for (int k = 0; k < Nq; ++k) {
    const int n = k * Nq2 + t;
    const T rx = g[0 * Np + n], ry = g[1 * Np + n], rz = g[2 * Np + n];
    const T sx = g[3 * Np + n], sy = g[4 * Np + n], sz = g[5 * Np + n];
    const T tx = g[6 * Np + n], ty = g[7 * Np + n], tz = g[8 * Np + n];
    const T JW = g[10 * Np + n];
    T grad[3][3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const T* s = s_q + f * Np;
      T a = 0, b = 0, c = 0;
#pragma unroll
      for (int m = 0; m < Nq; ++m) {
        a += s_D[i * Nq + m] * s[k * Nq2 + j * Nq + m];
        b += s_D[j * Nq + m] * s[k * Nq2 + m * Nq + i];
        c += s_D[k * Nq + m] * s[m * Nq2 + j * Nq + i];
      }
      grad[f][0] = rx * a + sx * b + tx * c;
      grad[f][1] = ry * a + sy * b + ty * c;
      grad[f][2] = rz * a + sz * b + tz * c;
    }
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const T lam0 = kLambdaField ? lambda0[(size_t)e * Np + n + (size_t)f * loffset] : lam0c[f];
      const T c0 = lam0 * JW;
      const T s1 = c0 * (grad[f][0] + grad[0][f]);
      const T s2 = c0 * (grad[f][1] + grad[1][f]);
      const T s3 = c0 * (grad[f][2] + grad[2][f]);
      s_F[(3 * f + 0) * Np + n] = rx * s1 + ry * s2 + rz * s3;
      s_F[(3 * f + 1) * Np + n] = sx * s1 + sy * s2 + sz * s3;
      s_F[(3 * f + 2) * Np + n] = tx * s1 + ty * s2 + tz * s3;
    }
  }
  __syncthreads();

  // sweep 2: weak divergence + mass term
#pragma unroll 1
  for (int k = 0; k < Nq; ++k) {
    const int n = k * Nq2 + t;
    const T JW = g[10 * Np + n];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const T *Fr = s_F + (3 * f + 0) * Np, *Fs = s_F + (3 * f + 1) * Np, *Ft = s_F + (3 * f + 2) * Np;
      T acc = 0;
#pragma unroll
      for (int m = 0; m < Nq; ++m) {
        acc += s_D[m * Nq + i] * Fr[k * Nq2 + j * Nq + m];
        acc += s_D[m * Nq + j] * Fs[k * Nq2 + m * Nq + i];
        acc += s_D[m * Nq + k] * Ft[m * Nq2 + j * Nq + i];
      }
      const T lam1 = lambda1[kLambdaField ? (size_t)e * Np + n + (size_t)f * loffset : (size_t)f * loffset];
      Aq[(size_t)e * Np + n + (size_t)f * offset] = acc + lam1 * JW * s_q[f * Np + n];
    }
  }
}

template <typename T, int Nq>
int launch_stress(dlong Nelements, dlong offset, dlong loffset, const dlong* elementList, const T* vgeo,
                  const T* D_host, const T* lambda0, const T* lambda1, int lambdaField, const T* q, T* Aq,
                  cudaStream_t stream)
{
  DMat<T, Nq> Dm;
  for (int n = 0; n < Nq * Nq; ++n) Dm.v[n] = D_host[n];
  const size_t smem = (size_t)12 * Nq * Nq * Nq * sizeof(T);
  if (lambdaField) {
    auto k = ax_stress_kernel<T, Nq, true>;
    if (smem > 40 * 1024) NRSB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<Nelements, Nq * Nq, smem, stream>>>(Nelements, offset, loffset, elementList, vgeo, Dm, lambda0, lambda1, q, Aq);
  } else {
    auto k = ax_stress_kernel<T, Nq, false>;
    if (smem > 40 * 1024) NRSB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<Nelements, Nq * Nq, smem, stream>>>(Nelements, offset, loffset, elementList, vgeo, Dm, lambda0, lambda1, q, Aq);
  }
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

}  // namespace

template <typename T>
int ax_stress_launch(int Nq, dlong Nelements, dlong offset, dlong loffset, const dlong* elementList, const T* vgeo,
                     const T* D_host, const T* lambda0, const T* lambda1, int lambdaField, const T* q, T* Aq,
                     cudaStream_t stream)
{
  if (Nelements == 0) return NRSB_OK;
#define CALL(n)                                                                                                 \
  case n:                                                                                                       \
    return launch_stress<T, n>(Nelements, offset, loffset, elementList, vgeo, D_host, lambda0, lambda1, lambdaField, \
                               q, Aq, stream);
  switch (Nq) {
    CALL(2)
    CALL(3)
    CALL(4)
    CALL(5)
    CALL(6)
    CALL(7)
    CALL(8)
    CALL(9)
    CALL(10)
    CALL(11)
    CALL(12)
    default: break;
  }
#undef CALL
  set_last_error("stress operator: unsupported Nq (supported: 2..12)");
  return NRSB_ERR_INVALID;
}

template int ax_stress_launch<double>(int, dlong, dlong, dlong, const dlong*, const double*, const double*,
                                      const double*, const double*, int, const double*, double*, cudaStream_t);
template int ax_stress_launch<float>(int, dlong, dlong, dlong, const dlong*, const float*, const float*, const float*,
                                     const float*, int, const float*, float*, cudaStream_t);

}  // namespace nrsb

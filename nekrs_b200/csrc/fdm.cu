// fdm.cu -- overlapping-Schwarz local solves by fast diagonalisation (fp32).
//
// Replaces kernels/elliptic/{preFDM,fusedFDM,postFDM}.okl (serial twins preFDM.c:15-93,
// fusedFDM.c:3-224, postFDM.c:14-153), called from pMGLevel::smoothSchwarz
// (ellipticMultiGridSchwarz.cpp:1056-1156).
//
//   u_e (extended (Nq+2)^3 element)  <-  (Sz x Sy x Sx) invL (Sz^T x Sy^T x Sx^T) u_e
//
// with "S v" = sum_l S[.][l] v[l] and "S^T v" = sum_l S[l][.] v[l] on the row-major per-element
// matrices the reference stores (gen_operators, ellipticMultiGridSchwarz.cpp:563-641).
//
// Kernel structure: one extended element per Nqe x Nqe thread slab, both work arrays in shared
// memory.  Each of the six contractions is executed by the thread that owns the whole pencil
// along the contracted direction: the pencil is read into registers once and the Nqe outputs are
// produced from it, so the per-element S matrices are the only shared-memory operand of the FMAs
// (broadcast reads).
#include "common.cuh"

namespace nrsb {

template <int Nqe>
struct FdmSmem {
  static constexpr int Np = Nqe * Nqe * Nqe;
  float A[Np];
  float B[Np];
  float Sx[Nqe * Nqe];
  float Sy[Nqe * Nqe];
  float Sz[Nqe * Nqe];
};

// out[o] = sum_l S(l,o) in[l]   (transposed = false: S[l*Nqe+o];  true: S[o*Nqe+l])
template <int Nqe, bool kTransposedS>
__device__ __forceinline__ void contract(const float* __restrict__ S, const float* in, int inStride, float* out,
                                         int outStride, const float* __restrict__ scale, int scaleStride)
{
  float v[Nqe];
#pragma unroll
  for (int l = 0; l < Nqe; ++l) v[l] = in[l * inStride];
#pragma unroll
  for (int o = 0; o < Nqe; ++o) {
    float acc = 0.f;
#pragma unroll
    for (int l = 0; l < Nqe; ++l) acc += (kTransposedS ? S[o * Nqe + l] : S[l * Nqe + o]) * v[l];
    if (scale) acc *= scale[o * scaleStride];
    out[o * outStride] = acc;
  }
}

template <int Nqe, bool kRestrict, int EPB>
__global__ void __launch_bounds__(Nqe* Nqe* EPB)
    fused_fdm_kernel(const dlong Nelements, const dlong* __restrict__ elementList, float* __restrict__ Su,
                     const float* __restrict__ S_x, const float* __restrict__ S_y, const float* __restrict__ S_z,
                     const float* __restrict__ inv_L, const float* __restrict__ wts, float* __restrict__ u)
{
  constexpr int Nq = Nqe - 2;
  constexpr int Nqe2 = Nqe * Nqe;
  constexpr int Npe = Nqe2 * Nqe;
  __shared__ FdmSmem<Nqe> sm[EPB];

  const int tid = threadIdx.x;
  const int t = tid % Nqe2;
  const int a = t % Nqe;
  const int b = t / Nqe;
  const int es = tid / Nqe2;
  const dlong e = blockIdx.x * EPB + es;
  const bool active = e < Nelements;
  const dlong element = active ? elementList[e] : 0;
  FdmSmem<Nqe>& s = sm[es];

  const float* ue = u + (size_t)element * Npe;
#pragma unroll
  for (int k = 0; k < Nqe; ++k) s.A[k * Nqe2 + t] = active ? ue[k * Nqe2 + t] : 0.f;
  s.Sx[t] = active ? S_x[(size_t)element * Nqe2 + t] : 0.f;
  s.Sy[t] = active ? S_y[(size_t)element * Nqe2 + t] : 0.f;
  s.Sz[t] = active ? S_z[(size_t)element * Nqe2 + t] : 0.f;
  __syncthreads();

  // remove the element's own contribution from the overlap planes (fusedFDM.c:33-92):
  // plane 0 -= plane 2, plane Nqe-1 -= plane Nqe-3, in each direction, interior of the face only
  if (a >= 1 && a < Nqe - 1 && b >= 1 && b < Nqe - 1) {
#define AIDX(k, j, i) ((k)*Nqe2 + (j)*Nqe + (i))
    s.A[AIDX(0, b, a)] -= s.A[AIDX(2, b, a)];
    s.A[AIDX(Nqe - 1, b, a)] -= s.A[AIDX(Nqe - 3, b, a)];
    s.A[AIDX(b, 0, a)] -= s.A[AIDX(b, 2, a)];
    s.A[AIDX(b, Nqe - 1, a)] -= s.A[AIDX(b, Nqe - 3, a)];
    s.A[AIDX(b, a, 0)] -= s.A[AIDX(b, a, 2)];
    s.A[AIDX(b, a, Nqe - 1)] -= s.A[AIDX(b, a, Nqe - 3)];
  }
  __syncthreads();

  // forward: S^T in x, y, z ; scale by invL
  contract<Nqe, false>(s.Sx, &s.A[AIDX(b, a, 0)], 1, &s.B[AIDX(b, a, 0)], 1, nullptr, 0);  // thread (j=a,k=b)
  __syncthreads();
  contract<Nqe, false>(s.Sy, &s.B[AIDX(b, 0, a)], Nqe, &s.A[AIDX(b, 0, a)], Nqe, nullptr, 0);  // (i=a,k=b)
  __syncthreads();
  contract<Nqe, false>(s.Sz, &s.A[AIDX(0, b, a)], Nqe2, &s.B[AIDX(0, b, a)], Nqe2,
                       inv_L + (size_t)element * Npe + t, Nqe2);  // (i=a,j=b)
  __syncthreads();
  // backward: S in x, y, z
  contract<Nqe, true>(s.Sx, &s.B[AIDX(b, a, 0)], 1, &s.A[AIDX(b, a, 0)], 1, nullptr, 0);
  __syncthreads();
  contract<Nqe, true>(s.Sy, &s.A[AIDX(b, 0, a)], Nqe, &s.B[AIDX(b, 0, a)], Nqe, nullptr, 0);
  __syncthreads();
  contract<Nqe, true>(s.Sz, &s.B[AIDX(0, b, a)], Nqe2, &s.A[AIDX(0, b, a)], Nqe2, nullptr, 0);
  // thread (i=a,j=b) now holds column (.,b,a) of the solution in s.A: no barrier needed to read it back

  if (!active) return;
  if (kRestrict) {
    // RAS: interior nodes only, weighted (fusedFDM.c:208-218)
    if (a >= 1 && a <= Nq && b >= 1 && b <= Nq) {
      const size_t base = (size_t)element * Nq * Nq * Nq + (b - 1) * Nq + (a - 1);
#pragma unroll
      for (int k = 0; k < Nq; ++k) {
        const size_t idx = base + (size_t)k * Nq * Nq;
        Su[idx] = s.A[AIDX(k + 1, b, a)] * wts[idx];
      }
    }
  } else {
    // ASM: full extended solution to Su; the overlap planes also go back to u (fusedFDM.c:160-206).
    // (The serial reference leaves an unspecified permuted intermediate in the rest of u; only
    //  the planes are ever read again, by postFDM.)
    float* Se = Su + (size_t)element * Npe;
    float* uo = u + (size_t)element * Npe;
#pragma unroll
    for (int k = 0; k < Nqe; ++k) {
      const float v = s.A[AIDX(k, b, a)];
      Se[k * Nqe2 + t] = v;
      uo[k * Nqe2 + t] = v;
    }
  }
#undef AIDX
}

// ---- preFDM: interior copy + overlap planes <- plane 2 of the element (preFDM.c:15-93)
template <int Nqe>
__global__ void __launch_bounds__(Nqe* Nqe) pre_fdm_kernel(const dlong Nelements, const float* __restrict__ u,
                                                          float* __restrict__ work1)
{
  constexpr int Nq = Nqe - 2;
  constexpr int Nqe2 = Nqe * Nqe;
  const dlong e = blockIdx.x;
  const int t = threadIdx.x;
  const int i = t % Nqe, j = t / Nqe;
  const float* ue = u + (size_t)e * Nq * Nq * Nq;
  float* w = work1 + (size_t)e * Nqe2 * Nqe;
#define UIDX(k, j, i) (((k)-1) * Nq * Nq + ((j)-1) * Nq + ((i)-1))
  const bool iin = i >= 1 && i < Nqe - 1, jin = j >= 1 && j < Nqe - 1;
#pragma unroll
  for (int k = 0; k < Nqe; ++k) {
    const bool kin = k >= 1 && k < Nqe - 1;
    float v = 0.f;
    if (iin && jin && kin)
      v = ue[UIDX(k, j, i)];
    else if (iin && jin && k == 0)
      v = ue[UIDX(2, j, i)];
    else if (iin && jin && k == Nqe - 1)
      v = ue[UIDX(Nqe - 3, j, i)];
    else if (iin && kin && j == 0)
      v = ue[UIDX(k, 2, i)];
    else if (iin && kin && j == Nqe - 1)
      v = ue[UIDX(k, Nqe - 3, i)];
    else if (jin && kin && i == 0)
      v = ue[UIDX(k, j, 2)];
    else if (jin && kin && i == Nqe - 1)
      v = ue[UIDX(k, j, Nqe - 3)];
    w[k * Nqe2 + t] = v;
  }
#undef UIDX
}

// ---- postFDM (ASM): fold the overlap back and weight (postFDM.c:14-153)
template <int Nqe>
__global__ void __launch_bounds__(Nqe* Nqe) post_fdm_kernel(const dlong Nelements, const float* __restrict__ my_work1,
                                                           const float* __restrict__ my_work2,
                                                           float* __restrict__ Su, const float* __restrict__ wts)
{
  constexpr int Nq = Nqe - 2;
  constexpr int Nqe2 = Nqe * Nqe;
  constexpr int Npe = Nqe2 * Nqe;
  __shared__ float w1[Npe];
  const dlong e = blockIdx.x;
  const int t = threadIdx.x;
  const int a = t % Nqe, b = t / Nqe;
  const float* g2 = my_work2 + (size_t)e * Npe;  // gathered extended solution
  const float* g1 = my_work1 + (size_t)e * Npe;  // this element's own overlap planes
#define AIDX(k, j, i) ((k)*Nqe2 + (j)*Nqe + (i))
#pragma unroll
  for (int k = 0; k < Nqe; ++k) w1[k * Nqe2 + t] = g2[k * Nqe2 + t];
  __syncthreads();
  const bool in = a >= 1 && a < Nqe - 1 && b >= 1 && b < Nqe - 1;
  if (in) {
    w1[AIDX(0, b, a)] -= g1[AIDX(0, b, a)];
    w1[AIDX(Nqe - 1, b, a)] -= g1[AIDX(Nqe - 1, b, a)];
    w1[AIDX(b, 0, a)] -= g1[AIDX(b, 0, a)];
    w1[AIDX(b, Nqe - 1, a)] -= g1[AIDX(b, Nqe - 1, a)];
    w1[AIDX(b, a, 0)] -= g1[AIDX(b, a, 0)];
    w1[AIDX(b, a, Nqe - 1)] -= g1[AIDX(b, a, Nqe - 1)];
  }
  __syncthreads();
  // the three folds are sequential in the reference (postFDM.c:90-140): plane 2 of one direction
  // may be read by the fold of the next direction
  if (in) {
    w1[AIDX(2, b, a)] += w1[AIDX(0, b, a)];
    w1[AIDX(Nqe - 3, b, a)] += w1[AIDX(Nqe - 1, b, a)];
  }
  __syncthreads();
  if (in) {
    w1[AIDX(b, 2, a)] += w1[AIDX(b, 0, a)];
    w1[AIDX(b, Nqe - 3, a)] += w1[AIDX(b, Nqe - 1, a)];
  }
  __syncthreads();
  if (in) {
    w1[AIDX(b, a, 2)] += w1[AIDX(b, a, 0)];
    w1[AIDX(b, a, Nqe - 3)] += w1[AIDX(b, a, Nqe - 1)];
  }
  __syncthreads();
  if (in) {
    const size_t base = (size_t)e * Nq * Nq * Nq + (b - 1) * Nq + (a - 1);
#pragma unroll
    for (int k = 0; k < Nq; ++k) {
      const size_t idx = base + (size_t)k * Nq * Nq;
      Su[idx] = w1[AIDX(k + 1, b, a)] * wts[idx];
    }
  }
#undef AIDX
}

// ------------------------------------------------------------------------------------------
template <int Nqe>
static int fused_launch(int restrict_, dlong Nelements, const dlong* elementList, float* Su, const float* Sx,
                        const float* Sy, const float* Sz, const float* invL, const float* wts, float* u,
                        cudaStream_t stream)
{
  constexpr int EPB = (Nqe * Nqe >= 64) ? 1 : (64 + Nqe * Nqe - 1) / (Nqe * Nqe);
  const int grid = (Nelements + EPB - 1) / EPB;
  if (restrict_)
    fused_fdm_kernel<Nqe, true, EPB><<<grid, Nqe * Nqe * EPB, 0, stream>>>(Nelements, elementList, Su, Sx, Sy, Sz,
                                                                            invL, wts, u);
  else
    fused_fdm_kernel<Nqe, false, EPB><<<grid, Nqe * Nqe * EPB, 0, stream>>>(Nelements, elementList, Su, Sx, Sy, Sz,
                                                                             invL, wts, u);
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

#define NRSB_FDM_SWITCH(CALL)                                       \
  switch (Nq + 2) {                                                 \
    case 4: CALL(4)                                                 \
    case 5: CALL(5)                                                 \
    case 6: CALL(6)                                                 \
    case 7: CALL(7)                                                 \
    case 8: CALL(8)                                                 \
    case 9: CALL(9)                                                 \
    case 10: CALL(10)                                               \
    case 11: CALL(11)                                               \
    case 12: CALL(12)                                               \
    case 13: CALL(13)                                               \
    case 14: CALL(14)                                               \
    default:                                                        \
      set_last_error("FDM: unsupported Nq (supported: 2..12)");     \
      return NRSB_ERR_INVALID;                                      \
  }

int fused_fdm_v1_launch(int Nq, int restrict_, dlong Nelements, const dlong* elementList, float* Su, const float* Sx,
                        const float* Sy, const float* Sz, const float* invL, const float* wts, float* u,
                        cudaStream_t stream, int epb, int warp);
bool fdm_supported(int Nq) { return Nq + 2 >= 4 && Nq + 2 <= 14; }
// 0: one pencil per thread (this file); 1: four pencils per thread (fdm_v1.cu), several elements per block;
// 2: the same with one element per warp where an element has <= 32 pencil owners (Nq = 8);  1x / 2x: developer
// override of the elements (warps) per block
static int g_fdm_variant = 2;
void set_fdm_variant(int v) { g_fdm_variant = v; }

int fused_fdm_launch(int Nq, int restrict_, dlong Nelements, const dlong* elementList, float* Su, const float* Sx,
                     const float* Sy, const float* Sz, const float* invL, const float* wts, float* u,
                     cudaStream_t stream)
{
  if (Nelements == 0) return NRSB_OK;
  if (g_fdm_variant >= 1) {
    const int warp = g_fdm_variant == 2 || g_fdm_variant >= 20;
    const int epb = g_fdm_variant >= 20 ? g_fdm_variant - 20 : (g_fdm_variant >= 10 ? g_fdm_variant - 10 : 0);
    const int rc = fused_fdm_v1_launch(Nq, restrict_, Nelements, elementList, Su, Sx, Sy, Sz, invL, wts, u, stream,
                                       epb, warp);
    if (rc != 1) return rc;  // 1 = size not covered (odd extended size): fall through to variant 0
  }
#define CALL(n) return fused_launch<n>(restrict_, Nelements, elementList, Su, Sx, Sy, Sz, invL, wts, u, stream);
  NRSB_FDM_SWITCH(CALL)
#undef CALL
}

int pre_fdm_launch(int Nq, dlong Nelements, const float* u, float* work1, cudaStream_t stream)
{
  if (Nelements == 0) return NRSB_OK;
#define CALL(n)                                                              \
  pre_fdm_kernel<n><<<Nelements, n * n, 0, stream>>>(Nelements, u, work1);   \
  break;
  NRSB_FDM_SWITCH(CALL)
#undef CALL
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

int post_fdm_launch(int Nq, dlong Nelements, const float* work1, const float* work2, float* Su, const float* wts,
                    cudaStream_t stream)
{
  if (Nelements == 0) return NRSB_OK;
#define CALL(n)                                                                           \
  post_fdm_kernel<n><<<Nelements, n * n, 0, stream>>>(Nelements, work1, work2, Su, wts);  \
  break;
  NRSB_FDM_SWITCH(CALL)
#undef CALL
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

}  // namespace nrsb

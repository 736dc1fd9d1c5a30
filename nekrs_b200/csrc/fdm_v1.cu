// fdm_v1.cu -- register-blocked fusedFDM for even extended sizes (Nqe = 4, 6, 8, 10, 12 <-> N = 1,3,5,7,9).
//
// fusedFDM is near the FP32 ridge (9.2 flop/B at N=7: 121 kflop vs 13.2 kB per element), so the kernel
// must feed the FMA pipe, not only HBM.  Variant 0 (fdm.cu) issues one broadcast shared-memory load of
// an S entry per FMA and sits on the LSU.  Here:
//   * every thread owns FOUR pencils (rows t, t+T, t+2T, t+3T of the [Nqe^2][Nqe] slab); a row of the
//     per-element S (padded to a multiple of 4) is fetched with 128-bit broadcast loads and reused for all
//     four: ~13 FMAs per S load instead of 1;
//   * every pass contracts the FASTEST index and writes its result TRANSPOSED (out[o][row]), so after
//     three passes the slab is back in natural [z][y][x] order and ALL six passes are the same code:
//     64-bit row loads whose lane stride (Nqe floats) is bank-conflict free for Nqe = 6, 10, and
//     unit-stride scalar stores;
//   * S and S^T are both staged so that forward and backward passes read contiguous rows;
//   * inverse eigenvalues are prefetched into registers before the first pass.
// Same contraction order as variant 0 (sum over l ascending) => identical results up to FMA contraction.
#include "common.cuh"

namespace nrsb {

namespace {

template <int Nqe>
struct FdmV1 {
  static constexpr int PPT = 4;                           // pencils per thread
  static constexpr int SP = (Nqe + 3) / 4 * 4;            // padded S row (floats)
  static constexpr int Nrows = Nqe * Nqe;
  static constexpr int Npe = Nrows * Nqe;
  static constexpr int TPE = Nrows / PPT;                 // threads per element
  static constexpr int smemFloats = 2 * Npe + 6 * Nqe * SP;
  // one-element-per-warp form: as many whole elements as fit the 32 lanes (Nqe = 10: 1, 8: 2, 6: 3, 4: 8); their
  // slabs are 8 floats apart beyond their size so that the elements of a warp start on different banks
  static constexpr int EPW = TPE <= 32 ? 32 / TPE : 1;
  static constexpr int warpStride = smemFloats + (EPW > 1 ? 8 : 0);
};

// two fp32 FMAs in one issue slot (Blackwell FFMA2; each half is an ordinary fma.rn, so the results are those of the
// scalar code).  The kernel is issue-bound (ncu: 58 % of the issue slots busy at 4 warps per scheduler, 70 % of the
// instructions FFMA): pairing the FMAs removes 35 % of the instructions.
__device__ __forceinline__ void ffma2(float2& acc, const float2 a, const float2 b)
{
  unsigned long long& c = reinterpret_cast<unsigned long long&>(acc);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(c)
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

// one contraction pass: out[o][row] = (scale) * sum_l S[l][o] in[row][l]   for the thread's PPT rows
template <int Nqe, bool kScale>
__device__ __forceinline__ void fdm_pass(const float* __restrict__ S, const float* in, float* out, int t,
                                         const float (&scale)[FdmV1<Nqe>::PPT][Nqe])
{
  using F = FdmV1<Nqe>;
  static_assert(Nqe % 2 == 0, "outputs are accumulated in pairs");
  float2 acc[F::PPT][Nqe / 2];
#pragma unroll
  for (int q = 0; q < F::PPT; ++q)
#pragma unroll
    for (int o = 0; o < Nqe / 2; ++o) acc[q][o] = make_float2(0.f, 0.f);
#pragma unroll
  for (int l = 0; l < Nqe; l += 2) {
    float2 v[F::PPT];
#pragma unroll
    for (int q = 0; q < F::PPT; ++q) v[q] = *reinterpret_cast<const float2*>(in + (t + F::TPE * q) * Nqe + l);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float2 s[F::SP / 2];
#pragma unroll
      for (int c = 0; c < F::SP / 4; ++c) {
        const float4 w = *reinterpret_cast<const float4*>(S + (l + h) * F::SP + 4 * c);
        s[2 * c + 0] = make_float2(w.x, w.y);
        s[2 * c + 1] = make_float2(w.z, w.w);
      }
#pragma unroll
      for (int q = 0; q < F::PPT; ++q) {
        const float x = h ? v[q].y : v[q].x;
        const float2 xx = make_float2(x, x);
#pragma unroll
        for (int o = 0; o < Nqe / 2; ++o) ffma2(acc[q][o], s[o], xx);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < Nqe; ++o)
#pragma unroll
    for (int q = 0; q < F::PPT; ++q) {
      float r = (o & 1) ? acc[q][o / 2].y : acc[q][o / 2].x;
      if (kScale) r *= scale[q][o];
      out[o * F::Nrows + t + F::TPE * q] = r;
    }
}

// kWarp (TPE <= 32): one element per WARP (lanes >= TPE idle).  The element's slabs are then private to the warp:
// no lane of another element shares a wavefront with it (the 10 880-byte slabs of Nqe = 10 all start on bank 0, and
// with 25 threads per element every warp straddled two of them: 39 % of the shared-memory wavefronts were conflicts,
// profiles/r2_ncu_fused_fdm_v1_E4096.md), the seven block barriers become __syncwarp, and the warps of a block drift
// apart so that one warp's global loads overlap another's contraction passes.
template <int Nqe, bool kRestrict, int EPB, bool kWarp>
__global__ void __launch_bounds__((kWarp ? 32 : FdmV1<Nqe>::TPE) * EPB, (kWarp && EPB <= 16) ? 16 / EPB : 1)
    fused_fdm_v1_kernel(const dlong Nelements, const dlong* __restrict__ elementList, float* __restrict__ Su,
                        const float* __restrict__ S_x, const float* __restrict__ S_y, const float* __restrict__ S_z,
                        const float* __restrict__ inv_L, const float* __restrict__ wts, float* __restrict__ u)
{
  using F = FdmV1<Nqe>;
  constexpr int Nq = Nqe - 2;
  constexpr int Nqe2 = Nqe * Nqe;
  constexpr int Npe = F::Npe;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  // kWarp: EPB warps per block, each with EPW elements of TPE lanes; else EPB elements of TPE threads
  constexpr int EPW = kWarp ? F::EPW : 1;
  const int t = kWarp ? (tid % 32) % F::TPE : tid % F::TPE;
  const int es = kWarp ? (tid / 32) * EPW + (tid % 32) / F::TPE : tid / F::TPE;
  const dlong e = blockIdx.x * (EPB * EPW) + es;
  const bool lane = !kWarp || (tid % 32) < EPW * F::TPE;  // this thread owns pencils
  const bool active = e < Nelements && lane;
  const dlong element = active ? elementList[e] : 0;
  auto sync = [] {
    if constexpr (kWarp)
      __syncwarp();
    else
      __syncthreads();
  };
  float* A = smem + (size_t)es * (kWarp ? F::warpStride : F::smemFloats);
  float* B = A + Npe;
  float* Sxf = B + Npe;  // forward  (row l: S[l][o])
  float* Syf = Sxf + Nqe * F::SP;
  float* Szf = Syf + Nqe * F::SP;
  float* Sxt = Szf + Nqe * F::SP;  // transposed (row l: S[o][l])
  float* Syt = Sxt + Nqe * F::SP;
  float* Szt = Syt + Nqe * F::SP;

  // ---- inverse eigenvalues of the entries this thread produces in pass 3 (natural index o*Nrows + row)
  float il[F::PPT][Nqe];
#pragma unroll
  for (int q = 0; q < F::PPT; ++q)
#pragma unroll
    for (int o = 0; o < Nqe; ++o)
      il[q][o] = active ? __ldcs(inv_L + (size_t)element * Npe + o * F::Nrows + t + F::TPE * q) : 0.f;

  // ---- stage the extended element (natural layout, 16-byte copies) and the S matrices (+ transposes).
  // All global loads of a thread are issued back to back into registers before the first dependent
  // shared-memory store: one DRAM round trip per element instead of one per loop iteration.
  {
    constexpr int NU = (Npe / 4 + F::TPE - 1) / F::TPE;
    constexpr int NS = (Nqe2 + F::TPE - 1) / F::TPE;
    const float4* src = reinterpret_cast<const float4*>(u + (size_t)element * Npe);
    float4 ru[NU];
    float rs[3][NS];
#pragma unroll
    for (int n = 0; n < NU; ++n) {
      const int idx = t + n * F::TPE;
      ru[n] = (active && idx < Npe / 4) ? __ldcs(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      const int idx = t + n * F::TPE;
      const bool ok = active && idx < Nqe2;
      rs[0][n] = ok ? __ldcs(S_x + (size_t)element * Nqe2 + idx) : 0.f;
      rs[1][n] = ok ? __ldcs(S_y + (size_t)element * Nqe2 + idx) : 0.f;
      rs[2][n] = ok ? __ldcs(S_z + (size_t)element * Nqe2 + idx) : 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(A);
#pragma unroll
    for (int n = 0; n < NU; ++n) {
      const int idx = t + n * F::TPE;
      if (lane && idx < Npe / 4) dst[idx] = ru[n];
    }
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      const int idx = t + n * F::TPE;
      if (lane && idx < Nqe2) {
        const int l = idx / Nqe, o = idx - l * Nqe;
        Sxf[l * F::SP + o] = rs[0][n];
        Syf[l * F::SP + o] = rs[1][n];
        Szf[l * F::SP + o] = rs[2][n];
        Sxt[o * F::SP + l] = rs[0][n];
        Syt[o * F::SP + l] = rs[1][n];
        Szt[o * F::SP + l] = rs[2][n];
      }
    }
  }
  sync();

  // ---- subtract the element's own contribution from the overlap planes (fusedFDM.c:33-92)
#define AI(k, j, i) (((k)*Nqe + (j)) * Nqe + (i))
  for (int idx = lane ? t : Nq * Nq; idx < Nq * Nq; idx += F::TPE) {
    const int a = 1 + idx % Nq, b = 1 + idx / Nq;
    A[AI(0, b, a)] -= A[AI(2, b, a)];
    A[AI(Nqe - 1, b, a)] -= A[AI(Nqe - 3, b, a)];
    A[AI(b, 0, a)] -= A[AI(b, 2, a)];
    A[AI(b, Nqe - 1, a)] -= A[AI(b, Nqe - 3, a)];
    A[AI(b, a, 0)] -= A[AI(b, a, 2)];
    A[AI(b, a, Nqe - 1)] -= A[AI(b, a, Nqe - 3)];
  }
  sync();

  // forward: S^T in x, y, z (each pass rotates the layout; three passes restore [z][y][x]), then scale
  if (lane) fdm_pass<Nqe, false>(Sxf, A, B, t, il);
  sync();
  if (lane) fdm_pass<Nqe, false>(Syf, B, A, t, il);
  sync();
  if (lane) fdm_pass<Nqe, true>(Szf, A, B, t, il);
  sync();
  // backward: S in x, y, z
  if (lane) fdm_pass<Nqe, false>(Sxt, B, A, t, il);
  sync();
  if (lane) fdm_pass<Nqe, false>(Syt, A, B, t, il);
  sync();
  // RAS weights of this thread's outputs: issued before the last pass so the latency hides behind it
  constexpr int NW = (Nq * Nq * Nq + F::TPE - 1) / F::TPE;
  float rw[kRestrict ? NW : 1];
  if (kRestrict) {
    const size_t base = (size_t)element * Nq * Nq * Nq;
#pragma unroll
    for (int n = 0; n < NW; ++n) {
      const int idx = t + n * F::TPE;
      rw[n] = (active && idx < Nq * Nq * Nq) ? __ldg(wts + base + idx) : 0.f;
    }
  }
  if (lane) fdm_pass<Nqe, false>(Szt, B, A, t, il);
  sync();

  if (!active) return;
  if (kRestrict) {
    // RAS: interior nodes, weighted (fusedFDM.c:208-218); coalesced over the element's Nq^3 outputs
    const size_t base = (size_t)element * Nq * Nq * Nq;
#pragma unroll
    for (int n = 0; n < NW; ++n) {
      const int idx = t + n * F::TPE;
      if (idx < Nq * Nq * Nq) {
        const int i = idx % Nq, j = (idx / Nq) % Nq, k = idx / (Nq * Nq);
        __stcs(Su + base + idx, A[AI(k + 1, j + 1, i + 1)] * rw[n]);
      }
    }
  } else {
    // ASM: extended solution to Su and back to u (the overlap planes are what postFDM reads)
    float4* Se = reinterpret_cast<float4*>(Su + (size_t)element * Npe);
    float4* uo = reinterpret_cast<float4*>(u + (size_t)element * Npe);
    const float4* src = reinterpret_cast<const float4*>(A);
    for (int idx = t; idx < Npe / 4; idx += F::TPE) {
      const float4 v = src[idx];
      Se[idx] = v;
      uo[idx] = v;
    }
  }
#undef AI
}

template <int Nqe, int EPBO = 0, bool kWarp = false>
int launch_v1(int restrict_, dlong Nelements, const dlong* elementList, float* Su, const float* Sx, const float* Sy,
              const float* Sz, const float* invL, const float* wts, float* u, cudaStream_t stream)
{
  using F = FdmV1<Nqe>;
  static_assert(!kWarp || F::TPE <= 32, "one element per warp needs at most 32 pencil owners");
  constexpr int want = (192 + F::TPE - 1) / F::TPE;
  constexpr int EPB = EPBO ? EPBO : (want > 32 ? 32 : (want < 1 ? 1 : want));
  constexpr int TPB = kWarp ? 32 : F::TPE;
  constexpr int EPW = kWarp ? F::EPW : 1;
  const size_t smem = (size_t)EPB * EPW * (kWarp ? F::warpStride : F::smemFloats) * sizeof(float);
  const int grid = (Nelements + EPB * EPW - 1) / (EPB * EPW);
  static bool configured[2] = {false, false};
  if (restrict_) {
    auto k = fused_fdm_v1_kernel<Nqe, true, EPB, kWarp>;
    if (!configured[1]) {
      NRSB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[1] = true;
    }
    k<<<grid, TPB * EPB, smem, stream>>>(Nelements, elementList, Su, Sx, Sy, Sz, invL, wts, u);
  } else {
    auto k = fused_fdm_v1_kernel<Nqe, false, EPB, kWarp>;
    if (!configured[0]) {
      NRSB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[0] = true;
    }
    k<<<grid, TPB * EPB, smem, stream>>>(Nelements, elementList, Su, Sx, Sy, Sz, invL, wts, u);
  }
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

}  // namespace

// returns 1 if this size is not handled here (odd Nqe).  epb > 0: developer override of the elements per block;
// warp: one element per warp (Nqe = 10: the default; epb = warps per block)
int fused_fdm_v1_launch(int Nq, int restrict_, dlong Nelements, const dlong* elementList, float* Su, const float* Sx,
                        const float* Sy, const float* Sz, const float* invL, const float* wts, float* u,
                        cudaStream_t stream, int epb, int warp)
{
#define ARGS restrict_, Nelements, elementList, Su, Sx, Sy, Sz, invL, wts, u, stream
  if (Nq == 8 && warp) {
    switch (epb) {
      case 2: return launch_v1<10, 2, true>(ARGS);
      case 8: return launch_v1<10, 8, true>(ARGS);
      default: return launch_v1<10, 4, true>(ARGS);
    }
  }
  if (warp && epb == 0) {  // the smaller even sizes: several elements per warp
    switch (Nq + 2) {
      case 4: return launch_v1<4, 4, true>(ARGS);
      case 6: return launch_v1<6, 4, true>(ARGS);
      case 8: return launch_v1<8, 4, true>(ARGS);
      default: break;
    }
  }
  if (Nq == 8 && epb) {
    switch (epb) {
      case 2: return launch_v1<10, 2>(ARGS);
      case 4: return launch_v1<10, 4>(ARGS);
      default: break;
    }
  }
  switch (Nq + 2) {
    case 4: return launch_v1<4>(ARGS);
    case 6: return launch_v1<6>(ARGS);
    case 8: return launch_v1<8>(ARGS);
    case 10: return launch_v1<10>(ARGS);
    case 12: return launch_v1<12>(ARGS);
    default: return 1;
  }
#undef ARGS
}

}  // namespace nrsb

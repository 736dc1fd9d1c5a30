// axhelm_tma.cu -- persistent, warp-specialised axhelm for Nq = 8 (N = 7): the element slabs
// (q: 1 plane set, ggeo: 6 or 7 planes) are streamed into a shared-memory ring by ONE producer
// lane with bulk async copies (cp.async.bulk ... mbarrier::complete_tx, the TMA engine's 1-D path),
// while groups of Nq^2 consumer threads each work on one element with the pencil algorithm of
// axhelm.inc (variant 1).  One CTA per SM, grid = #SMs: per-SM bytes in flight are set by the ring
// depth (NSTAGES x 28 KB), not by occupancy, and the tail is at most one element per group.
//
// Why: at E = 4096 the non-persistent kernels lose ~30 % to wave quantisation (4096 blocks over
// 148 x 8 slots) and to the load -> compute -> store phases of each short-lived block; the ring
// keeps HBM busy through all phases (ncu of variant 1: DRAM 43 %, occupancy-limited).
//
// Same arithmetic order as variant 1 => identical results.
#include <cstdlib>

#include "common.cuh"
#include "gs.hpp"
#include "halo.cuh"
#include "kernels.hpp"

namespace nrsb {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// geometric factors are read exactly once per operator application: mark them evict-first so that the
// 117 MB stream does not push the E-vectors (q, Aq: what the gather-scatter touches next) out of L2
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void group_sync(int id, int nthreads)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- phase 2 of the kGs launch: this thread's share of the on-rank gather rows and masked nodes.
// Thread `gt` of `nT`; rows are dealt round-robin so that neighbouring lanes read neighbouring table entries.
// The table entries of the first batch (kGsP pair rows + kGsQ quad rows per thread: everything a 4096-element
// box needs) are fetched BEFORE the device-wide barrier, so that after it one round of value loads (L2, __ldcg:
// other SMs stored them) and the stores remain.  Copies are summed in ascending local index, the reference's
// order (gatherScatterMany.okl), so the result is bit-identical to the separate kernel's.
constexpr int kGsP = 20;
constexpr int kGsQ = 4;
struct GsPrefetch {
  int2 p[kGsP];
  int4 q[kGsQ];
};

__device__ __forceinline__ void gs_phase_prefetch(const GsRowsDev& R, const int gt, const int nT, GsPrefetch& F)
{
#pragma unroll
  for (int j = 0; j < kGsP; ++j) {
    const int i = gt + j * nT;
    F.p[j] = i < R.nPairs ? __ldg(R.pairs + i) : make_int2(-1, -1);
  }
#pragma unroll
  for (int j = 0; j < kGsQ; ++j) {
    const int i = gt + j * nT;
    F.q[j] = i < R.nQuads ? __ldg(R.quads + i) : make_int4(-1, -1, -1, -1);
  }
}

template <typename T>
__device__ __forceinline__ void gs_phase_rows(const GsRowsDev& R, const int gt, const int nT, T* __restrict__ q,
                                              GsPrefetch& F)
{
  for (int it = 0; (long)it * kGsP * nT < R.nPairs || (long)it * kGsQ * nT < R.nQuads; ++it) {
    if (it > 0) {  // later batches (larger meshes): same code, entries fetched here
#pragma unroll
      for (int j = 0; j < kGsP; ++j) {
        const long i = gt + ((long)it * kGsP + j) * nT;
        F.p[j] = i < R.nPairs ? __ldg(R.pairs + i) : make_int2(-1, -1);
      }
#pragma unroll
      for (int j = 0; j < kGsQ; ++j) {
        const long i = gt + ((long)it * kGsQ + j) * nT;
        F.q[j] = i < R.nQuads ? __ldg(R.quads + i) : make_int4(-1, -1, -1, -1);
      }
    }
    T a[kGsP], b[kGsP], v[kGsQ][4];
#pragma unroll
    for (int j = 0; j < kGsP; ++j)
      if (F.p[j].x >= 0) {
        a[j] = __ldcg(q + F.p[j].x);
        b[j] = __ldcg(q + F.p[j].y);
      }
#pragma unroll
    for (int j = 0; j < kGsQ; ++j)
      if (F.q[j].x >= 0) {
        v[j][0] = __ldcg(q + F.q[j].x);
        v[j][1] = __ldcg(q + F.q[j].y);
        v[j][2] = __ldcg(q + F.q[j].z);
        v[j][3] = __ldcg(q + F.q[j].w);
      }
#pragma unroll
    for (int j = 0; j < kGsP; ++j)
      if (F.p[j].x >= 0) {
        T sum = T(0);
        sum += a[j];
        sum += b[j];
        q[F.p[j].x] = sum;
        q[F.p[j].y] = sum;
      }
#pragma unroll
    for (int j = 0; j < kGsQ; ++j)
      if (F.q[j].x >= 0) {
        T sum = T(0);
        sum += v[j][0];
        sum += v[j][1];
        sum += v[j][2];
        sum += v[j][3];
        q[F.q[j].x] = sum;
        q[F.q[j].y] = sum;
        q[F.q[j].z] = sum;
        q[F.q[j].w] = sum;
      }
  }
  for (int i = gt; i < R.nOcts; i += nT) {
    const int4 ia = __ldg(R.octs + 2 * i), ib = __ldg(R.octs + 2 * i + 1);
    const T v0 = __ldcg(q + ia.x), v1 = __ldcg(q + ia.y), v2 = __ldcg(q + ia.z), v3 = __ldcg(q + ia.w);
    const T v4 = __ldcg(q + ib.x), v5 = __ldcg(q + ib.y), v6 = __ldcg(q + ib.z), v7 = __ldcg(q + ib.w);
    T sum = T(0);
    sum += v0;
    sum += v1;
    sum += v2;
    sum += v3;
    sum += v4;
    sum += v5;
    sum += v6;
    sum += v7;
    q[ia.x] = sum;
    q[ia.y] = sum;
    q[ia.z] = sum;
    q[ia.w] = sum;
    q[ib.x] = sum;
    q[ib.y] = sum;
    q[ib.z] = sum;
    q[ib.w] = sum;
  }
  for (int i = gt; i < R.nGen; i += nT) {
    const int s0 = __ldg(R.genStarts + i), s1 = __ldg(R.genStarts + i + 1);
    T sum = T(0);
    for (int c = s0; c < s1; ++c) sum += __ldcg(q + __ldg(R.genIds + c));
    for (int c = s0; c < s1; ++c) q[__ldg(R.genIds + c)] = sum;
  }
  for (int i = gt; i < R.nMasked; i += nT) q[__ldg(R.maskIds + i)] = T(0);
}

template <typename T, int Nq>
struct SlabT {
  static constexpr int bankMod = 32 / (sizeof(T) / 4);
  static constexpr int plane()
  {
    int p = Nq * Nq;
    while (p % bankMod != Nq % bankMod) ++p;
    return p;
  }
  static constexpr int P = plane();
  static constexpr int size = P * Nq;
  __device__ static __forceinline__ int rot(int i, int j)
  {
    int r = i + j;
    return r >= Nq ? r - Nq : r;
  }
  __device__ static __forceinline__ int idx(int i, int j, int k) { return rot(i, j) + Nq * j + P * k; }
};

}  // namespace

// kFused: the first F.NhaloElements entries of the element list are the elements that touch another rank.
// The last F.nPush CTAs of the grid take no elements: they wait until all halo elements are stored
// (device-scope counter), then pack the halo rows and push the partial sums into the neighbours' receive
// windows over NVLink while the other CTAs carry on with the interior elements: ellipticOperator's
// "Ax(halo) -> oogs::start -> Ax(interior)" (ellipticOperator.cpp:117-172) in ONE launch.  All CTAs are
// co-resident (grid = #SMs, one CTA per SM), so the wait cannot deadlock; it is bounded anyway.
// (Measured: a service warp inside a working CTA needs ~20 us for the push, each dependent access queues
// behind that SM's 200 KB of in-flight bulk copies; dedicated CTAs need ~5 us.)
//
// kGs: ellipticOperator's mask + on-rank gather-scatter (ellipticOperator.cpp:158-168) become phase 2 of the same
// launch (struct FusedRows, gs.hpp): no kernel boundary, no second launch, row tables fetched while waiting.
template <typename T, int Nq, int NGROUPS, int NSTAGES, bool kPoisson, bool kFused, bool kGs, bool kDot>
__global__ void __launch_bounds__(NGROUPS* Nq* Nq + 32, 1)
    ax_tma_kernel(const dlong Nelements, const dlong* __restrict__ elementList, const T* __restrict__ ggeo,
                  const DMat<T, Nq> Dm, const T* __restrict__ lambda0, const T* __restrict__ lambda1,
                  const T* __restrict__ q, T* __restrict__ Aq, const FusedHalo F, const FusedRows R2)
{
  constexpr int Np = Nq * Nq * Nq;
  constexpr int Nq2 = Nq * Nq;
  constexpr int NG = kPoisson ? 6 : 7;  // geometric-factor planes needed (G00..G22 [, GwJ]); contiguous
  constexpr uint32_t gBytes = NG * Np * sizeof(T);
  constexpr uint32_t qBytes = Np * sizeof(T);
  constexpr int stageElems = (NG + 1) * Np;
  static_assert(NSTAGES % NGROUPS == 0, "each consumer group must own a private sub-ring");
  using S = SlabT<T, Nq>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* stages = reinterpret_cast<T*>(smem_raw);                        // [NSTAGES][(NG+1)*Np]
  T* work = stages + (size_t)NSTAGES * stageElems;                   // [NGROUPS][3][S::size]
  uint64_t* full = reinterpret_cast<uint64_t*>(work + (size_t)NGROUPS * 3 * S::size);
  uint64_t* empty = full + NSTAGES;

  const int tid = threadIdx.x;
  constexpr int nConsumers = NGROUPS * Nq2;
  // the gather-scatter launch that follows on the stream (gs.cu / oogs.cu) may become resident next to this
  // CTA right away: it fetches its index tables and then sleeps in griddepcontrol.wait until this grid is done
  pdl_trigger();
  // kFused: the last F.nPush CTAs do no element work at all, they are the halo pushers (a CTA whose SM is
  // saturated by the TMA ring pays microseconds per dependent load; an otherwise idle SM does not)
  const int nAx = kFused ? (int)gridDim.x - F.nPush : (int)gridDim.x;
  const bool pusher = kFused && (int)blockIdx.x >= nAx;
  const int myCount =
      (!pusher && Nelements > (dlong)blockIdx.x) ? (int)((Nelements - blockIdx.x + nAx - 1) / nAx) : 0;

  if (tid == 0) {
    for (int s = 0; s < NSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], Nq2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (pusher) {
    if (kDot && tid == 0) R2.dotPartials[blockIdx.x] = 0.0;
    // ===== halo pusher CTA: all threads =====
    const HaloExchangeDev& H = F.H;
    const int pb = blockIdx.x - nAx;
    // the send table does not depend on the results: fetch the first batch while the halo elements are
    // still being computed
    // the send table does not depend on the results: fetch this thread's entries while the halo elements
    // are still being computed.  Under a streaming load every dependent access costs 2-3 us (measured), so
    // after the wait only  value load -> NVLink store -> fence -> flag  remains on the critical path.
    const int estride = F.nPush * blockDim.x;
    const int e00 = pb * blockDim.x + tid;
    HaloSendBatch<16> first;
    halo_pack_load<16>(H, e00, estride, first);
    const bool stamp = F.stamps && pb == 0 && tid == 0;
    if (stamp) F.stamps[0] = globaltimer_ns();
    if (tid == 0) {
      const long long t0 = clock64();
      while (ld_acquire_u64(F.counter) < F.target) {
        __nanosleep(64);
        if (clock64() - t0 > (1ll << 33)) {  // ~4 s: co-residency of the grid was violated (MPS / MIG / a smaller part)
          if (F.err) *F.err = 1;               // the host turns this into NRSB_ERR_CUDA at its next synchronisation
          break;
        }
      }
    }
    __syncthreads();
    if (stamp) F.stamps[1] = globaltimer_ns();
    {
      T val[16];
      halo_pack_gather<T, 16, true>(H, gs_op::add, Aq, e00, estride, first, val);
      if (stamp) F.stamps[5] = globaltimer_ns() + (val[0] == T(12345.678) ? 1 : 0);  // (after the values arrived)
      halo_pack_scatter<T, 16>(H, (T*)F.partial, e00, estride, first, val);
    }
    for (int e0 = e00 + 16 * estride; e0 < H.nSend; e0 += 8 * estride)
      halo_pack_flat<T, 8, true>(H, gs_op::add, Aq, (T*)F.partial, e0, estride);
    if (stamp) F.stamps[2] = globaltimer_ns();
    __threadfence_system();
    __syncthreads();
    if (stamp) F.stamps[3] = globaltimer_ns();
    // this pusher's rows are out: raise ITS flag slot at every peer (receivers wait for all kFlagSlots slots)
    for (int i = tid; i < H.nPeers * kFlagSlots; i += blockDim.x) {
      const int p = i / kFlagSlots, slot = i % kFlagSlots;
      if (slot % F.nPush != pb) continue;
      volatile unsigned long long* f = H.peerFlags[p] + (size_t)H.myRank * kFlagSlots + slot;
      *f = H.epoch;
    }
    if (stamp) F.stamps[4] = globaltimer_ns();
    return;
  }
  if (tid >= nConsumers) {
    // ===== producer warp: one lane streams element slabs into the ring =====
    if (tid == nConsumers) {
      const uint64_t pol = l2_evict_first_policy();
      for (int i = 0; i < myCount; ++i) {
        const int s = i % NSTAGES;
        if (i >= NSTAGES) mbar_wait(&empty[s], ((i / NSTAGES) - 1) & 1);
        const dlong element = elementList[blockIdx.x + (dlong)i * nAx];
        T* st = stages + (size_t)s * stageElems;
        mbar_expect_tx(&full[s], gBytes + qBytes);
        bulk_g2s_hint(st, ggeo + (size_t)element * 7 * Np, gBytes, &full[s], pol);
        bulk_g2s(st + NG * Np, q + (size_t)element * Np, qBytes, &full[s]);
      }
    }
    return;
  }

  // ===== consumers: group g owns elements g, g+NGROUPS, ... of this CTA =====
  const int g = tid / Nq2;
  const int t = tid % Nq2;
  const int a = t % Nq;
  const int b = t / Nq;
  T* su = work + (size_t)g * 3 * S::size;
  T* sr = su + S::size;
  T* ss = sr + S::size;
  const int tbase = S::rot(a, b) + Nq * b;
  const T lam0 = lambda0[0];
  const T lam1 = kPoisson ? T(0) : lambda1[0];
  T dotAcc = T(0);  // kDot: this thread's share of q^T A q

  for (int i = g; i < myCount; i += NGROUPS) {
    const int s = i % NSTAGES;
    const dlong element = elementList[blockIdx.x + (dlong)i * nAx];
    const T* st = stages + (size_t)s * stageElems;
    const T* sq = st + NG * Np;
    mbar_wait(&full[s], (i / NSTAGES) & 1);

    // 1. q in the t-layout + swizzled copy
    T r_q[Nq];
#pragma unroll
    for (int k = 0; k < Nq; ++k) r_q[k] = sq[t + Nq2 * k];
#pragma unroll
    for (int k = 0; k < Nq; ++k) su[tbase + S::P * k] = r_q[k];
    group_sync(1 + g, Nq2);

    // 2. derivatives along the owned pencils
    T r_qt[Nq];
#pragma unroll
    for (int k = 0; k < Nq; ++k) {
      T v = 0;
#pragma unroll
      for (int m = 0; m < Nq; ++m) v += Dm.v[k * Nq + m] * r_q[m];
      r_qt[k] = v;
    }
    {
      T u[Nq];
#pragma unroll
      for (int m = 0; m < Nq; ++m) u[m] = su[S::idx(m, a, b)];
#pragma unroll
      for (int ii = 0; ii < Nq; ++ii) {
        T v = 0;
#pragma unroll
        for (int m = 0; m < Nq; ++m) v += Dm.v[ii * Nq + m] * u[m];
        sr[S::idx(ii, a, b)] = v;
      }
#pragma unroll
      for (int m = 0; m < Nq; ++m) u[m] = su[S::idx(a, m, b)];
#pragma unroll
      for (int jj = 0; jj < Nq; ++jj) {
        T v = 0;
#pragma unroll
        for (int m = 0; m < Nq; ++m) v += Dm.v[jj * Nq + m] * u[m];
        ss[S::idx(a, jj, b)] = v;
      }
    }
    group_sync(1 + g, Nq2);

    // 3. geometric factors from the staged slab (t-layout: conflict-free)
    T r_mass[kPoisson ? 1 : Nq];
#pragma unroll
    for (int k = 0; k < Nq; ++k) {
      const int n = t + Nq2 * k;
      const T G00 = st[0 * Np + n], G01 = st[1 * Np + n], G11 = st[2 * Np + n];
      const T G12 = st[3 * Np + n], G02 = st[4 * Np + n], G22 = st[5 * Np + n];
      if constexpr (!kPoisson) r_mass[k] = st[6 * Np + n] * lam1 * r_q[k];
      const int p = tbase + S::P * k;
      const T qr = sr[p], qs = ss[p], qt = r_qt[k];
      T Gqr = G00 * qr;
      Gqr += G01 * qs;
      Gqr += G02 * qt;
      T Gqs = G01 * qr;
      Gqs += G11 * qs;
      Gqs += G12 * qt;
      T Gqt = G02 * qr;
      Gqt += G12 * qs;
      Gqt += G22 * qt;
      if constexpr (kDot) {
        T e = qr * Gqr;
        e += qs * Gqs;
        e += qt * Gqt;
        if constexpr (!kPoisson) e = lam0 * e + r_mass[k] * r_q[k];
        dotAcc += e;
      }
      sr[p] = lam0 * Gqr;
      ss[p] = lam0 * Gqs;
      r_qt[k] = lam0 * Gqt;
      // kDot: keep ptxas from hoisting all 64 shared loads of this loop above the first FMA (it then spills)
      if constexpr (kDot)
        if (k == Nq / 2 - 1) __syncwarp();
    }
    // the stage is no longer needed: hand it back to the producer
    mbar_arrive(&empty[s]);
    group_sync(1 + g, Nq2);

    // 4. transposed derivatives
    T r_Aq[Nq];
#pragma unroll
    for (int k = 0; k < Nq; ++k) {
      T v = 0;
#pragma unroll
      for (int m = 0; m < Nq; ++m) v += Dm.v[m * Nq + k] * r_qt[m];
      r_Aq[k] = v;
    }
    {
      T u[Nq];
#pragma unroll
      for (int m = 0; m < Nq; ++m) u[m] = sr[S::idx(m, a, b)];
#pragma unroll
      for (int ii = 0; ii < Nq; ++ii) {
        T v = 0;
#pragma unroll
        for (int m = 0; m < Nq; ++m) v += Dm.v[m * Nq + ii] * u[m];
        su[S::idx(ii, a, b)] = v;
      }
#pragma unroll
      for (int m = 0; m < Nq; ++m) u[m] = ss[S::idx(a, m, b)];
#pragma unroll
      for (int jj = 0; jj < Nq; ++jj) {
        T v = 0;
#pragma unroll
        for (int m = 0; m < Nq; ++m) v += Dm.v[m * Nq + jj] * u[m];
        ss[S::idx(a, jj, b)] = v;
      }
    }
    group_sync(1 + g, Nq2);

    // 5. sum and store (coalesced)
    T* Ae = Aq + (size_t)element * Np + t;
#pragma unroll
    for (int k = 0; k < Nq; ++k) {
      const int p = tbase + S::P * k;
      T v = r_Aq[k] + su[p] + ss[p];
      if constexpr (!kPoisson) v += r_mass[k];
      Ae[k * Nq2] = v;
    }
    const bool haloElem = kFused && (blockIdx.x + (dlong)i * nAx < F.NhaloElements);
    group_sync(1 + g, Nq2);  // su/ss are rewritten by the next element of this group
    if (haloElem && t == 0) {
      // the group's stores are ordered before this point by the barrier; one fence (cumulative) publishes them
      __threadfence();
      atomicAdd(F.counter, 1ull);
    }
  }

  if (kFused && F.stamps && tid == 0) {
    if (blockIdx.x == 0) F.stamps[8] = globaltimer_ns();
    if ((int)blockIdx.x == nAx - 1) F.stamps[9] = globaltimer_ns();
  }
  if constexpr (kDot) {
    // fixed-order fold: lanes (shuffle tree) -> warps (ascending) -> one partial per CTA
    double* s_dot = reinterpret_cast<double*>(empty + NSTAGES);
    const double wsum = warp_sum(kPoisson ? (double)(lam0 * dotAcc) : (double)dotAcc);
    if ((tid & 31) == 0) s_dot[tid >> 5] = wsum;
    group_sync(14, nConsumers);
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < nConsumers / 32; ++w) tot += s_dot[w];
      R2.dotPartials[blockIdx.x] = tot;
    }
  }

  if constexpr (kGs) {
    // ===== phase 2: mask + on-rank gather-scatter by the consumer threads of all axhelm CTAs =====
    const int nT = nAx * nConsumers;
    const int gt = blockIdx.x * nConsumers + tid;  // a warp reads 32 consecutive table entries
    GsPrefetch pre;
    gs_phase_prefetch(R2.R, gt, nT, pre);
    group_sync(15, nConsumers);
    if (tid == 0) {
      __threadfence();  // the CTA's stores (ordered before the barrier above) become visible device-wide
      atomicAdd(R2.arrive, 1ull);
      const long long t0 = clock64();
      while (ld_acquire_u64(R2.arrive) < R2.target) {
        __nanosleep(32);
        if (clock64() - t0 > (1ll << 33)) {  // ~4 s: co-residency of the grid was violated
          if (R2.err) *R2.err = 1;
          break;
        }
      }
    }
    group_sync(15, nConsumers);
    gs_phase_rows<T>(R2.R, gt, nT, Aq, pre);
  }
}

template <typename T, int Nq, int NGROUPS, int NSTAGES, bool kPoisson, bool kFused = false, bool kGs = false,
          bool kDot = false>
static int launch_tma(dlong Nelements, const dlong* elementList, const T* ggeo, const T* D_host, const T* lambda0,
                      const T* lambda1, const T* q, T* Aq, cudaStream_t stream, const FusedHalo* fused = nullptr,
                      FusedRows* rows = nullptr, AxDot* dot = nullptr)
{
  constexpr int Np = Nq * Nq * Nq;
  constexpr int NG = kPoisson ? 6 : 7;
  using S = SlabT<T, Nq>;
  const size_t smem = ((size_t)NSTAGES * (NG + 1) * Np + (size_t)NGROUPS * 3 * S::size) * sizeof(T) +
                      2 * NSTAGES * sizeof(uint64_t) + 128;
  auto kern = ax_tma_kernel<T, Nq, NGROUPS, NSTAGES, kPoisson, kFused, kGs, kDot>;
  static bool configured = false;
  if (!configured) {
    NRSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  DMat<T, Nq> Dm;
  for (int n = 0; n < Nq * Nq; ++n) Dm.v[n] = D_host[n];
  int grid = kNumSMs;
  if (grid > Nelements) grid = Nelements;
  const FusedHalo F = fused ? *fused : FusedHalo();
  if (kFused) grid = kNumSMs;  // pushers + workers, all co-resident
  FusedRows R2;
  if (kGs) {
    R2 = *rows;
    R2.target += (unsigned long long)(kFused ? grid - F.nPush : grid);  // every axhelm CTA arrives once
    rows->target = R2.target;
  }
  if (kDot) {
    R2.dotPartials = dot->partials;
    dot->n = grid;
  }
  // (Launching this kernel as a programmatic dependent launch, with the first geometric-factor slabs requested
  // before griddepcontrol.wait, was measured slower: 28-29 us per launch with the attribute, 31-33 us without it,
  // against 26.7 us for the plain launch -- the wait itself resolves late.)
  kern<<<grid, NGROUPS * Nq * Nq + 32, smem, stream>>>(Nelements, elementList, ggeo, Dm, lambda0, lambda1, q, Aq, F,
                                                       R2);
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

// variants 4..6 of the axhelm dispatch (Nq = 8, constant coefficients only)
template <typename T>
int ax_tma_launch(int Nq, int variant, dlong Nelements, const dlong* elementList, const T* ggeo, const T* D_host,
                  const T* lambda0, const T* lambda1, int poisson, const T* q, T* Aq, cudaStream_t stream, AxDot* dot)
{
  if (Nelements == 0) return NRSB_OK;
  if (Nq != 8) {
    set_last_error("axhelm variants 4-6 (TMA ring) are built for Nq = 8 only");
    return NRSB_ERR_INVALID;
  }
#define NRSB_TMA(G, S_)                                                                                              \
  return poisson ? launch_tma<T, 8, G, S_, true>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, q, Aq, stream, \
                                                 nullptr, nullptr, dot)                                                 \
                 : launch_tma<T, 8, G, S_, false>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, q, Aq, stream, \
                                                  nullptr, nullptr, dot);
  // NSTAGES must be a multiple of NGROUPS: group g then only ever touches stages == g (mod NGROUPS),
  // i.e. it owns a private sub-ring, and every mbarrier wait is at most one phase behind.
  // Helmholtz stages carry a seventh plane (GwJ): 6 x 32 KB + work buffers exceed the 227 KB of an SM for the
  // 3x6 / 5x5 (fp64) and 6x12 / 8x8 (fp32) rings, so the Helmholtz operator always takes the 4-stage ring.
  if (!poisson) variant = 4;
  if (sizeof(T) == 8) {
    if (variant == 4) { NRSB_TMA(4, 4) }
    if (variant == 5) {
      if (dot && dot->partials)
        return poisson ? launch_tma<T, 8, 3, 6, true, false, false, true>(Nelements, elementList, ggeo, D_host, lambda0,
                                                                          lambda1, q, Aq, stream, nullptr, nullptr, dot)
                       : launch_tma<T, 8, 3, 6, false, false, false, true>(Nelements, elementList, ggeo, D_host,
                                                                           lambda0, lambda1, q, Aq, stream, nullptr,
                                                                           nullptr, dot);
      NRSB_TMA(3, 6)
    }
    NRSB_TMA(5, 5)
  } else {
    if (variant == 4) { NRSB_TMA(4, 8) }
    if (variant == 5) { NRSB_TMA(6, 12) }
    NRSB_TMA(8, 8)
  }
#undef NRSB_TMA
}

// Ax over [halo elements..., interior elements...] with the halo push done inside the launch (kFused)
template <typename T>
int ax_tma_fused_launch(int Nq, int variant, dlong Nelements, const dlong* elementList, const T* ggeo, const T* D_host,
                        const T* lambda0, const T* lambda1, int poisson, const T* q, T* Aq, const FusedHalo& F,
                        cudaStream_t stream, AxDot* dot)
{
  if (Nelements == 0) return NRSB_OK;
  if (Nq != 8) {
    set_last_error("fused axhelm + halo push is built for Nq = 8 only");
    return NRSB_ERR_INVALID;
  }
  (void)variant;
#define NRSB_TMAF(G, S_)                                                                                      \
  return poisson ? launch_tma<T, 8, G, S_, true, true>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, q, \
                                                       Aq, stream, &F, nullptr, dot)                           \
                 : launch_tma<T, 8, G, S_, false, true>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, q, \
                                                        Aq, stream, &F, nullptr, dot);
  if (sizeof(T) == 8) {
    if (dot && dot->partials)
      return poisson ? launch_tma<T, 8, 3, 6, true, true, false, true>(Nelements, elementList, ggeo, D_host, lambda0,
                                                                       lambda1, q, Aq, stream, &F, nullptr, dot)
                     : launch_tma<T, 8, 3, 6, false, true, false, true>(Nelements, elementList, ggeo, D_host, lambda0,
                                                                        lambda1, q, Aq, stream, &F, nullptr, dot);
    NRSB_TMAF(3, 6)
  }
  NRSB_TMAF(6, 12)
#undef NRSB_TMAF
}
// Ax + mask + on-rank gather-scatter in ONE launch (single rank: the whole ellipticOperator; several ranks:
// halo rows are pushed by the pusher CTAs and folded by oogs::finish).  rows->target is advanced.
template <typename T>
int ax_tma_gs_launch(int Nq, dlong Nelements, const dlong* elementList, const T* ggeo, const T* D_host,
                     const T* lambda0, const T* lambda1, int poisson, const T* q, T* Aq, const FusedHalo* F,
                     FusedRows* rows, cudaStream_t stream, AxDot* dot)
{
  if (Nelements == 0) return NRSB_OK;
  if (Nq != 8) {
    set_last_error("fused axhelm + gather-scatter is built for Nq = 8 only");
    return NRSB_ERR_INVALID;
  }
#define NRSB_TMAG(G_, S_, FUSED)                                                                                   \
  return poisson ? launch_tma<T, 8, G_, S_, true, FUSED, true>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, \
                                                              q, Aq, stream, F, rows, dot)                           \
                 : launch_tma<T, 8, G_, S_, false, FUSED, true>(Nelements, elementList, ggeo, D_host, lambda0,       \
                                                               lambda1, q, Aq, stream, F, rows, dot);
  if (F) {
    if (sizeof(T) == 8) { NRSB_TMAG(3, 6, true) }
    NRSB_TMAG(6, 12, true)
  }
  if (sizeof(T) == 8) { NRSB_TMAG(3, 6, false) }
  NRSB_TMAG(6, 12, false)
#undef NRSB_TMAG
}
template int ax_tma_gs_launch<double>(int, dlong, const dlong*, const double*, const double*, const double*,
                                      const double*, int, const double*, double*, const FusedHalo*, FusedRows*,
                                      cudaStream_t, AxDot*);
template int ax_tma_gs_launch<float>(int, dlong, const dlong*, const float*, const float*, const float*, const float*,
                                     int, const float*, float*, const FusedHalo*, FusedRows*, cudaStream_t, AxDot*);

template int ax_tma_fused_launch<double>(int, int, dlong, const dlong*, const double*, const double*, const double*,
                                         const double*, int, const double*, double*, const FusedHalo&, cudaStream_t,
                                         AxDot*);
template int ax_tma_fused_launch<float>(int, int, dlong, const dlong*, const float*, const float*, const float*,
                                        const float*, int, const float*, float*, const FusedHalo&, cudaStream_t,
                                        AxDot*);

template int ax_tma_launch<double>(int, int, dlong, const dlong*, const double*, const double*, const double*,
                                   const double*, int, const double*, double*, cudaStream_t, AxDot*);
template int ax_tma_launch<float>(int, int, dlong, const dlong*, const float*, const float*, const float*,
                                  const float*, int, const float*, float*, cudaStream_t, AxDot*);

}  // namespace nrsb

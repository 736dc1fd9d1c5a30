// reduce.cuh -- single-launch reduction template shared by linalg.cu and coarse.cu.
//
// grid-stride accumulation in fp64 -> warp shuffle + shared-memory block fold -> per-block partials ->
// the LAST block (atomic ticket) folds the partials in a fixed order, performs the cross-GPU one-shot
// all-reduce through peer-mapped windows and runs an optional `post` functor on the totals (used to
// update Krylov scalars on the device without another launch).
#pragma once
#include "linalg.hpp"

namespace nrsb {

constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* s_red)
{
  // s_red: kRedThreads/32 doubles.  Result valid in thread 0.
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < kRedThreads / 32) ? s_red[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

// cross-rank one-shot all-reduce executed by WARP 0 of the last block: lane p talks to rank p, so the nranks
// NVLink transfers (store words into the peer's window; poll the own window for the peer's words) run side by side
// instead of one after the other.  The contributions are then added in ascending rank order by every lane
// (shuffles), i.e. the same bits on every rank.
// epoch e, parity-double-buffered slots: slots[p] = peer p's window [2][nranks][kMaxRed][2 words].
// `vals` (shared memory, nv entries) holds this rank's totals on entry and the global totals on exit.
template <int NV>
__device__ __forceinline__ void peer_allreduce_warp(const PeerReduce& P, double* vals, int nv)
{
  const int lane = threadIdx.x & 31;
  unsigned long long e = *P.epoch + 1ull;
  if ((e & 0xffffffffull) == 0ull) ++e;  // 0 is "never written"
  __syncwarp();
  if (lane == 0) *P.epoch = e;
  const int par = (int)(e & 1ull);
  const unsigned long long tag = (e & 0xffffffffull) << 32;
  double tot[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) tot[v] = 0.0;
  for (int base = 0; base < P.nranks; base += 32) {  // more than 32 ranks: batches of 32 peers
    const int p = base + lane;
    double c[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) c[v] = 0.0;
    if (p < P.nranks) {
      // every value travels WITH its flag: two 8-byte words {low half | epoch}, {high half | epoch}; an aligned
      // 8-byte store is single-copy atomic over NVLink, so no fence and no separate flag are needed (the
      // fence + flag version cost ~20 us per all-reduce at 2 GPUs: two system-scope fences on the critical path)
      volatile unsigned long long* dst =
          (volatile unsigned long long*)P.slots[p] + (((size_t)par * P.nranks + P.rank) * kMaxRed) * 2;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (v < nv) {
          const unsigned long long b = (unsigned long long)__double_as_longlong(vals[v]);
          dst[2 * v] = tag | (b & 0xffffffffull);
          dst[2 * v + 1] = tag | (b >> 32);
        }
      const volatile unsigned long long* loc =
          (const volatile unsigned long long*)P.slots[P.rank] + (((size_t)par * P.nranks + p) * kMaxRed) * 2;
      const long long t0 = clock64();
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (v < nv) {
          unsigned long long lo = loc[2 * v], hi = loc[2 * v + 1];
          while ((lo >> 32) != (tag >> 32) || (hi >> 32) != (tag >> 32)) {
            if (clock64() - t0 > (1ll << 34)) {  // ~8 s: a peer never arrived; flag it instead of hanging the box
              if (P.err) *P.err = 1;
              break;
            }
            lo = loc[2 * v];
            hi = loc[2 * v + 1];
          }
          c[v] = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
        }
    }
    __syncwarp();
    const int cnt = P.nranks - base < 32 ? P.nranks - base : 32;
    for (int q = 0; q < cnt; ++q) {  // ascending rank: same bits everywhere
#pragma unroll
      for (int v = 0; v < NV; ++v) tot[v] += __shfl_sync(0xffffffffu, c[v], q);
    }
  }
  if (lane == 0)
    for (int v = 0; v < NV; ++v)
      if (v < nv) vals[v] = tot[v];
  __syncwarp();
}

struct NoPost {
  __device__ __forceinline__ void operator()(double*) const {}
};

template <int NV, typename Op, typename Post>
__global__ void __launch_bounds__(kRedThreads)
    reduce_kernel(long N, Op op, int nv, double* out, ReduceWs ws, Post post)
{
  __shared__ double s_red[kRedThreads / 32];
  __shared__ double s_tot[NV];
  __shared__ bool s_last;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) op(i, acc);

#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double b = block_sum(acc[v], s_red);
    if (threadIdx.x == 0 && v < nv) ws.partials[(size_t)blockIdx.x * kMaxRed + v] = b;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(ws.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double tot[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double a = 0.0;
    if (v < nv)
      for (int b = threadIdx.x; b < (int)gridDim.x; b += kRedThreads)
        a += __ldcg(&ws.partials[(size_t)b * kMaxRed + v]);
    tot[v] = block_sum(a, s_red);
  }
  if (ws.peer.nranks > 1) {
    if (threadIdx.x == 0)
      for (int v = 0; v < NV; ++v) s_tot[v] = tot[v];
    __syncthreads();
    if (threadIdx.x < 32) peer_allreduce_warp<NV>(ws.peer, s_tot, nv);
    if (threadIdx.x == 0)
      for (int v = 0; v < NV; ++v) tot[v] = s_tot[v];
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (v < nv) out[v] = tot[v];
    post(tot);
    *ws.ticket = 0u;
  }
}

static inline int red_grid(long N, int perThread = 8)
{
  long b = (N + (long)kRedThreads * perThread - 1) / ((long)kRedThreads * perThread);
  if (b > kMaxRedBlocks) b = kMaxRedBlocks;
  if (b < 1) b = 1;
  return (int)b;
}

template <int NV, typename Op, typename Post = NoPost>
static int reduce_launch(long N, Op op, int nv, double* out, const ReduceWs& ws, cudaStream_t s, Post post = Post(),
                         int perThread = 8)
{
  if (!ws.partials || !ws.ticket) {
    set_last_error("reduction workspace not initialised");
    return NRSB_ERR_INVALID;
  }
  reduce_kernel<NV, Op, Post><<<red_grid(N, perThread), kRedThreads, 0, s>>>(N, op, nv, out, ws, post);
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

}  // namespace nrsb

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

// cross-rank one-shot all-reduce executed by thread 0 of the last block.
// epoch e, parity-double-buffered slots: slots[p] = peer p's window [2][nranks][kMaxRed]
__device__ __forceinline__ void peer_allreduce(const PeerReduce& P, double* vals, int nv)
{
  const unsigned long long e = *P.epoch + 1ull;
  *P.epoch = e;
  const int par = (int)(e & 1ull);
  for (int p = 0; p < P.nranks; ++p) {
    double* dst = P.slots[p] + ((size_t)par * P.nranks + P.rank) * kMaxRed;
    for (int v = 0; v < nv; ++v) dst[v] = vals[v];  // NVLink stores
  }
  __threadfence_system();
  for (int p = 0; p < P.nranks; ++p) {
    volatile unsigned long long* f = P.flags[p] + P.rank;
    *f = e;
  }
  volatile unsigned long long* mine = P.flags[P.rank];
  const volatile double* loc = P.slots[P.rank] + (size_t)par * P.nranks * kMaxRed;
  for (int v = 0; v < nv; ++v) vals[v] = 0.0;
  for (int p = 0; p < P.nranks; ++p) {
    while (mine[p] < e) {
    }
    __threadfence_system();
    for (int v = 0; v < nv; ++v) vals[v] += loc[(size_t)p * kMaxRed + v];  // ascending rank: same bits everywhere
  }
}

struct NoPost {
  __device__ __forceinline__ void operator()(double*) const {}
};

template <int NV, typename Op, typename Post>
__global__ void __launch_bounds__(kRedThreads)
    reduce_kernel(long N, Op op, int nv, double* out, ReduceWs ws, Post post)
{
  __shared__ double s_red[kRedThreads / 32];
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
  if (threadIdx.x == 0) {
    if (ws.peer.nranks > 1) peer_allreduce(ws.peer, tot, nv);
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

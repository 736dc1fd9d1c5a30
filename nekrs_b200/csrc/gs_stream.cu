// gs_stream.cu -- the on-rank gather-scatter + Dirichlet mask of ellipticOperator (ellipticOperator.cpp:158-168:
// ellipticApplyMask, oogs::startFinish(ogsAdd)) run CONCURRENTLY with the persistent axhelm launch that produces
// its input, instead of as a second pass after it.
//
// Why (B200, E = 4096, N = 7, fp64): axhelm 27 us + gather-scatter 12 us.  The second launch moves data that is
// entirely L2 resident, half of its time is kernel boundary + ramp, and it cannot start before the LAST element is
// stored.  But a row only needs ITS elements: 90 % of the rows are final long before axhelm ends.
//
// How: the axhelm CTA (255 registers x 224 threads, 214 KB shared memory) leaves 8192 registers and ~18 KB of shared
// memory per SM.  This kernel is one 128-thread block per SM at <= 64 registers and no shared memory, launched as
// a programmatic dependent launch, so its blocks become resident next to the axhelm CTAs as soon as all of those
// have started (griddepcontrol.launch_dependents at their top).  axhelm counts finished elements per chunk of
// consecutive element-list positions (fence + device-scope add after the element's stores); the warps here walk
// warp-sized units of rows sorted by ready chunk: table entries are fetched BEFORE the wait (they do not depend on
// the data), then one acquire poll of all chunk counters, one round of L2 loads (ld.global.cg: other SMs stored the
// values), the sums in ascending local index (the reference's order: bit-identical to gs_rows_kernel), the stores.
// If the blocks cannot be co-resident (MPS, profiler serialisation) the kernel simply runs after axhelm: the
// counters are already complete, every wait falls through.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "gs.hpp"

namespace nrsb {

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kThreads = 128;

// highest c such that chunks 0..c are all complete (-1: none).  Every lane acquires one counter; __syncwarp orders
// the other lanes' later loads after it (barrier synchronisation is part of the causality order).
__device__ __forceinline__ int chunks_ready(const GsStreamDev& S, const int lane)
{
  bool ok = true;
  if (lane < S.nChunks) {
    const long first = (long)lane * S.chunkLen;
    const long size = min((long)S.chunkLen, (long)S.Nelements - first);
    ok = ld_acquire_gpu(S.done + lane) >= S.epoch * (unsigned long long)size;
  }
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  __syncwarp();
  return (m == 0xffffffffu) ? 31 : __ffs(~m) - 2;
}

}  // namespace

template <typename T>
__global__ void __launch_bounds__(kThreads, 8)  // <= 64 registers: fits beside the 255-register axhelm CTA
    gs_stream_kernel(const GsStreamDev S, T* __restrict__ q)
{
  constexpr int RP = gs_stream_t::kPairsPerLane, RQ = gs_stream_t::kQuadsPerLane, RM = gs_stream_t::kMaskPerLane;
  static_assert(2 * RP == 12 && 4 * RQ == 12 && RM == 12, "12 node ids per lane");
  const int lane = threadIdx.x & 31;
  const int W = gridDim.x * (kThreads / 32);
  const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;  // consecutive units land on different SMs
  int readyUpTo = -1;
  for (int u = gw; u < S.nUnits; u += W) {
    const int4 d = __ldg(S.units + u);
    const int kind = d.x, chunk = d.y, first = d.z, count = d.w;
    if (kind == 4 && !S.withMask) continue;
    // ---- table entries first (independent of the data): 12 node ids per lane, whatever the row length
    //      (6 pairs, 3 quads, 1 octet [+4 unused], 12 masked nodes; general rows: CSR bounds in id[0], id[1])
    int id[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) id[i] = -1;
    if (kind == 0) {
#pragma unroll
      for (int j = 0; j < RP; ++j) {
        const int r = lane + 32 * j;
        if (r < count) {
          const int2 e = __ldg(S.pairs + first + r);
          id[2 * j] = e.x;
          id[2 * j + 1] = e.y;
        }
      }
    } else if (kind == 1) {
#pragma unroll
      for (int j = 0; j < RQ; ++j) {
        const int r = lane + 32 * j;
        if (r < count) {
          const int4 e = __ldg(S.quads + first + r);
          id[4 * j] = e.x;
          id[4 * j + 1] = e.y;
          id[4 * j + 2] = e.z;
          id[4 * j + 3] = e.w;
        }
      }
    } else if (kind == 2) {
      if (lane < count) {
        const int4 e0 = __ldg(S.octs + 2 * (first + lane)), e1 = __ldg(S.octs + 2 * (first + lane) + 1);
        id[0] = e0.x;
        id[1] = e0.y;
        id[2] = e0.z;
        id[3] = e0.w;
        id[4] = e1.x;
        id[5] = e1.y;
        id[6] = e1.z;
        id[7] = e1.w;
      }
    } else if (kind == 3) {
      if (lane < count) {
        id[0] = __ldg(S.genStarts + first + lane);
        id[1] = __ldg(S.genStarts + first + lane + 1);
      }
    } else {
#pragma unroll
      for (int j = 0; j < RM; ++j) {
        const int r = lane + 32 * j;
        if (r < count) id[j] = __ldg(S.maskIds + first + r);
      }
    }
    // ---- wait until every element of chunks 0..chunk is final in global memory
    if (chunk > readyUpTo) {
      const long long t0 = clock64();
      while ((readyUpTo = chunks_ready(S, lane)) < chunk) {
        __nanosleep(200);
        if (clock64() - t0 > (1ll << 33)) {  // ~4 s: the producer never ran (mis-use); flag it, do not hang
          if (lane == 0 && S.err) *S.err = 1;
          break;
        }
      }
    }
    // ---- values (L2), sums in ascending local index, stores
    if (kind <= 2) {
      T v[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) v[i] = id[i] >= 0 ? __ldcg(q + id[i]) : T(0);
      if (kind == 0) {
#pragma unroll
        for (int j = 0; j < RP; ++j) {
          T s = T(0);
          s += v[2 * j];
          s += v[2 * j + 1];
          v[2 * j] = s;
          v[2 * j + 1] = s;
        }
      } else if (kind == 1) {
#pragma unroll
        for (int j = 0; j < RQ; ++j) {
          T s = T(0);
          s += v[4 * j];
          s += v[4 * j + 1];
          s += v[4 * j + 2];
          s += v[4 * j + 3];
          v[4 * j] = v[4 * j + 1] = v[4 * j + 2] = v[4 * j + 3] = s;
        }
      } else {
        T s = T(0);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = s;
      }
#pragma unroll
      for (int i = 0; i < 12; ++i)
        if (id[i] >= 0) q[id[i]] = v[i];
    } else if (kind == 3) {
      if (lane < count) {
        T s = T(0);
        for (int c = id[0]; c < id[1]; ++c) s += __ldcg(q + __ldg(S.genIds + c));
        for (int c = id[0]; c < id[1]; ++c) q[__ldg(S.genIds + c)] = s;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i)
        if (id[i] >= 0) q[id[i]] = T(0);
    }
  }
  // the producer grid must be complete (and its other results, e.g. the q^T A q partials, visible) before this
  // grid completes: later kernels of the stream only wait for THIS grid
  pdl_wait();
}

template <typename T>
int gs_stream_launch(const GsStreamDev& S, T* q, cudaStream_t stream)
{
  if (S.nUnits == 0) return NRSB_OK;
  auto kern = gs_stream_kernel<T>;
  static bool configured = false;
  if (!configured) {
    // same shared-memory carve-out as the axhelm CTA it has to sit next to
    NRSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  NRSB_CUDA(launch_pdl_if(true, kern, dim3(kNumSMs), dim3(kThreads), 0, stream, S, q));
  return NRSB_OK;
}
template int gs_stream_launch<double>(const GsStreamDev&, double*, cudaStream_t);
template int gs_stream_launch<float>(const GsStreamDev&, float*, cudaStream_t);

bool gs_stream_fits(int axRegsPerThread, int axThreads, size_t axSmemBytes)
{
  cudaFuncAttributes fa{}, fb{};
  if (cudaFuncGetAttributes(&fa, gs_stream_kernel<double>) != cudaSuccess) return false;
  if (cudaFuncGetAttributes(&fb, gs_stream_kernel<float>) != cudaSuccess) return false;
  const int regs = std::max(fa.numRegs, fb.numRegs);
  auto warpRegs = [](int r) { return ((r + 7) / 8) * 8 * 32; };  // allocation unit: 8 registers per thread
  const long used = (long)warpRegs(axRegsPerThread) * ((axThreads + 31) / 32) + (long)warpRegs(regs) * (kThreads / 32);
  const size_t smem = axSmemBytes + 1024 + fa.sharedSizeBytes + 1024;
  return used <= 65536 && smem <= 233472 && axThreads + kThreads <= 2048;
}

// ------------------------------------------------------------------------------------------------ host tables
gs_stream_t::~gs_stream_t()
{
  cudaFree(d_pairs);
  cudaFree(d_quads);
  cudaFree(d_octs);
  cudaFree(d_genStarts);
  cudaFree(d_genIds);
  cudaFree(d_maskIds);
  cudaFree(d_units);
  cudaFree(d_done);
  if (h_err) cudaFreeHost(h_err);
}

namespace {
template <typename V>
int up(V** dst, const std::vector<V>& h)
{
  cudaFree(*dst);
  *dst = nullptr;
  if (h.empty()) return NRSB_OK;
  NRSB_CUDA(cudaMalloc((void**)dst, h.size() * sizeof(V)));
  NRSB_CUDA(cudaMemcpy(*dst, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice));
  return NRSB_OK;
}
}  // namespace

int gs_stream_t::build(const ogs_t* ogs, const std::vector<dlong>& maskIds, const std::vector<dlong>& elementPos,
                       int Np, int nAx)
{
  Nelements = (int)elementPos.size();
  NRSB_REQUIRE(Nelements > 0 && nAx > 0, "gs_stream: empty element list");
  // <= kMaxChunks chunks of whole axhelm rounds (nAx elements finish at about the same time), at least two rounds
  // each so that the per-chunk wait is amortised
  {
    const long rounds = (Nelements + nAx - 1) / nAx;
    long per = std::max(2l, (rounds + kMaxChunks - 1) / kMaxChunks);
    chunkLen = (int)(per * nAx);
    nChunks = (Nelements + chunkLen - 1) / chunkLen;
  }
  auto chunkOfNode = [&](dlong n) { return (int)(elementPos[n / Np] / chunkLen); };

  struct Row {
    int chunk;
    int idx;  // original order (ascending base id)
  };
  std::vector<std::vector<int>> P(nChunks), Q(nChunks), O(nChunks), G(nChunks), M(nChunks);
  for (dlong r = 0; r < ogs->NlocalGather; ++r) {
    const dlong s = ogs->localGatherOffsets[r], cnt = ogs->localGatherOffsets[r + 1] - s;
    if (cnt == 1) continue;
    int c = 0;
    for (dlong k = 0; k < cnt; ++k) c = std::max(c, chunkOfNode(ogs->localGatherIds[s + k]));
    (cnt == 2 ? P : cnt == 4 ? Q : cnt == 8 ? O : G)[c].push_back((int)r);
  }
  for (dlong n : maskIds) M[chunkOfNode(n)].push_back((int)n);

  std::vector<int2> pairs;
  std::vector<int4> quads, octs, units;
  std::vector<int> genStarts(1, 0), genIds, mask;
  auto cut = [&](int kind, int chunk, int first, int n, int perUnit) {
    for (int o = 0; o < n; o += perUnit) units.push_back(make_int4(kind, chunk, first + o, std::min(perUnit, n - o)));
  };
  for (int c = 0; c < nChunks; ++c) {
    const int p0 = (int)pairs.size(), q0 = (int)quads.size(), o0 = (int)octs.size() / 2, g0 = (int)genStarts.size() - 1,
              m0 = (int)mask.size();
    for (int r : P[c]) {
      const dlong* g = &ogs->localGatherIds[ogs->localGatherOffsets[r]];
      pairs.push_back(make_int2(g[0], g[1]));
    }
    for (int r : Q[c]) {
      const dlong* g = &ogs->localGatherIds[ogs->localGatherOffsets[r]];
      quads.push_back(make_int4(g[0], g[1], g[2], g[3]));
    }
    for (int r : O[c]) {
      const dlong* g = &ogs->localGatherIds[ogs->localGatherOffsets[r]];
      octs.push_back(make_int4(g[0], g[1], g[2], g[3]));
      octs.push_back(make_int4(g[4], g[5], g[6], g[7]));
    }
    for (int r : G[c]) {
      for (dlong k = ogs->localGatherOffsets[r]; k < ogs->localGatherOffsets[r + 1]; ++k)
        genIds.push_back(ogs->localGatherIds[k]);
      genStarts.push_back((int)genIds.size());
    }
    for (int n : M[c]) mask.push_back(n);
    cut(0, c, p0, (int)P[c].size(), 32 * kPairsPerLane);
    cut(1, c, q0, (int)Q[c].size(), 32 * kQuadsPerLane);
    cut(2, c, o0, (int)O[c].size(), 32);
    cut(3, c, g0, (int)G[c].size(), 32);
    cut(4, c, m0, (int)M[c].size(), 32 * kMaskPerLane);
  }
  nUnits = (int)units.size();
  int rc;
  if ((rc = up(&d_pairs, pairs))) return rc;
  if ((rc = up(&d_quads, quads))) return rc;
  if ((rc = up(&d_octs, octs))) return rc;
  if ((rc = up(&d_genStarts, genStarts))) return rc;
  if ((rc = up(&d_genIds, genIds))) return rc;
  if ((rc = up(&d_maskIds, mask))) return rc;
  if ((rc = up(&d_units, units))) return rc;
  if (!d_done) {
    NRSB_CUDA(cudaMalloc((void**)&d_done, kMaxChunks * sizeof(unsigned long long)));
  }
  NRSB_CUDA(cudaMemset(d_done, 0, kMaxChunks * sizeof(unsigned long long)));
  NRSB_CUDA(cudaDeviceSynchronize());
  epoch = 0;
  if (!h_err) {
    NRSB_CUDA(cudaHostAlloc((void**)&h_err, sizeof(int), cudaHostAllocMapped));
    *h_err = 0;
    NRSB_CUDA(cudaHostGetDevicePointer((void**)&d_err, h_err, 0));
  }
  return NRSB_OK;
}

GsStreamDev gs_stream_t::dev(bool withMask) const
{
  GsStreamDev S;
  S.nUnits = nUnits;
  S.units = d_units;
  S.pairs = d_pairs;
  S.quads = d_quads;
  S.octs = d_octs;
  S.genStarts = d_genStarts;
  S.genIds = d_genIds;
  S.maskIds = d_maskIds;
  S.nChunks = nChunks;
  S.chunkLen = chunkLen;
  S.Nelements = Nelements;
  S.done = d_done;
  S.epoch = epoch;
  S.withMask = withMask ? 1 : 0;
  S.err = d_err;
  return S;
}

}  // namespace nrsb

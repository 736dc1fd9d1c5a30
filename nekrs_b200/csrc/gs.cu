// gs.cu -- on-rank gather-scatter  v[i] <- sum_{copies j of i} v[j]   (+ Dirichlet mask).
//
// Replaces: 3rd_party/gslib/ogs/okl/gatherScatterMany.okl (gatherScatterMany_{float,double}Add),
// kernels/core/mask.okl, and the device side of the halo exchange, okl/oogs.okl packBuf/unpackBuf
// (oogs.okl:1-272), as called from oogs::start/finish (oogs.cpp:682-821).
//
// Layout (B200-first, results identical): the reference walks a CSR (gatherStarts, gatherIds)
// with one thread per row and skips singleton rows at run time (`start+1 != end`).  On a hex
// mesh almost every non-singleton row has 2 (face), 4 (edge) or 8 (vertex) copies, so at setup
// the rows are bucketed by length: pairs as int2, quads as int4, octets as 2 x int4, the rest as
// CSR.  A thread reads its row's indices with ONE coalesced vector load instead of two offset
// loads plus a dependent index loop, and singleton rows are not stored at all.  Inside a row the
// copies are summed in ascending local index exactly as the CSR loop does (ogsSetup.cpp:196-249
// ordering), so results are bit-identical to the reference's kernel.
//
// The mask (q[maskIds[n]] = 0) is folded into the same launch: masked nodes carry global id 0 in
// the masked ogs handle (ellipticOgs.cpp:126-131) and therefore belong to no row, so zeroing them
// commutes with the sums.
#include <cstdlib>

#include "common.cuh"
#include "gs.hpp"
#include "halo.cuh"

namespace nrsb {

// One block = one row kind (pairs / quads / octets / general / masked nodes), chosen by blockIdx.x, so a block has
// no divergence and every thread holds only what its kind needs: kRP pair rows per thread are 2 kRP index words and
// 2 kRP values, all loads of a thread issued before the first add (predicated, no branches).  The grid is sized to
// fit ONE wave of 2048 threads per SM (16 blocks of 128 threads at 32 registers): the second wave of the
// one-row-per-thread kernel cost a second round of  index load -> value load -> store  latencies (11.7 us -> see
// profiles/ for the E = 4096 numbers).  Copies are summed in ascending local index, as the reference's CSR loop does
// (bit-identical sums).
// The kernel is launched as a programmatic dependent launch: blocks may become resident while the producer
// (axhelm) is still running, fetch their index entries, and sleep in pdl_wait() until its stores are visible.
struct GsGrid {
  int pairBlocks, quadBlocks, octBlocks, genBlocks, maskBlocks;
};
constexpr int kGsBS = 128;
constexpr int kGsMaskPerThread = 4;

template <typename T, int kRP>
__device__ __forceinline__ void gs_rows_body(const GsRowsDev& R, const GsGrid& G, int b, T* __restrict__ qf)
{
  const int t = threadIdx.x;
  if (b < G.pairBlocks) {
    int2 id[kRP];
#pragma unroll
    for (int j = 0; j < kRP; ++j) {
      const int r = (b * kRP + j) * kGsBS + t;
      id[j] = r < R.nPairs ? R.pairs[r] : make_int2(-1, -1);
    }
    pdl_wait();
    T a[kRP], c[kRP];
#pragma unroll
    for (int j = 0; j < kRP; ++j)
      if (id[j].x >= 0) {
        a[j] = qf[id[j].x];
        c[j] = qf[id[j].y];
      }
#pragma unroll
    for (int j = 0; j < kRP; ++j)
      if (id[j].x >= 0) {
        T s = T(0);
        s += a[j];
        s += c[j];
        qf[id[j].x] = s;
        qf[id[j].y] = s;
      }
    return;
  }
  b -= G.pairBlocks;
  if (b < G.quadBlocks) {
    const int r = b * kGsBS + t;
    const int4 id = r < R.nQuads ? R.quads[r] : make_int4(-1, -1, -1, -1);
    pdl_wait();
    if (id.x < 0) return;
    const T v0 = qf[id.x], v1 = qf[id.y], v2 = qf[id.z], v3 = qf[id.w];
    T s = T(0);
    s += v0;
    s += v1;
    s += v2;
    s += v3;
    qf[id.x] = s;
    qf[id.y] = s;
    qf[id.z] = s;
    qf[id.w] = s;
    return;
  }
  b -= G.quadBlocks;
  if (b < G.octBlocks) {
    const int r = b * kGsBS + t;
    int4 ia = make_int4(-1, -1, -1, -1), ib = ia;
    if (r < R.nOcts) {
      ia = R.octs[2 * r];
      ib = R.octs[2 * r + 1];
    }
    pdl_wait();
    if (ia.x < 0) return;
    const T v0 = qf[ia.x], v1 = qf[ia.y], v2 = qf[ia.z], v3 = qf[ia.w];
    const T v4 = qf[ib.x], v5 = qf[ib.y], v6 = qf[ib.z], v7 = qf[ib.w];
    T s = T(0);
    s += v0;
    s += v1;
    s += v2;
    s += v3;
    s += v4;
    s += v5;
    s += v6;
    s += v7;
    qf[ia.x] = s;
    qf[ia.y] = s;
    qf[ia.z] = s;
    qf[ia.w] = s;
    qf[ib.x] = s;
    qf[ib.y] = s;
    qf[ib.z] = s;
    qf[ib.w] = s;
    return;
  }
  b -= G.octBlocks;
  if (b < G.genBlocks) {
    const int r = b * kGsBS + t;
    int s0 = 0, s1 = 0;
    if (r < R.nGen) {
      s0 = R.genStarts[r];
      s1 = R.genStarts[r + 1];
    }
    pdl_wait();
    T s = T(0);
    for (int c = s0; c < s1; ++c) s += qf[R.genIds[c]];
    for (int c = s0; c < s1; ++c) qf[R.genIds[c]] = s;
    return;
  }
  b -= G.genBlocks;
  {
    int id[kGsMaskPerThread];
#pragma unroll
    for (int j = 0; j < kGsMaskPerThread; ++j) {
      const int r = (b * kGsMaskPerThread + j) * kGsBS + t;
      id[j] = r < R.nMasked ? R.maskIds[r] : -1;
    }
    pdl_wait();
#pragma unroll
    for (int j = 0; j < kGsMaskPerThread; ++j)
      if (id[j] >= 0) qf[id[j]] = T(0);
  }
}

template <typename T, int kRP>
__global__ void __launch_bounds__(kGsBS, 2048 / kGsBS)
    gs_rows_kernel(const GsRowsDev R, const GsGrid G, const dlong stride, T* __restrict__ q)
{
  pdl_trigger();  // a persistent axhelm launch behind this kernel may start its prologue on drained SMs
  gs_rows_body<T, kRP>(R, G, blockIdx.x, q + (size_t)blockIdx.y * stride);
}

// oogs::finish for one field and ogsAdd (the operator's case): the on-rank rows and the mask exactly as above, plus
// one more block kind at the END of the grid for the halo rows: table entries first, then the wait for the peers'
// epoch flags (normally long raised: the peers pushed while this rank was still in axhelm), then
//   own partial + received partials in ascending rank order -> every local copy
// (unpackBuf of okl/oogs.okl:121-272 + the scatter).  Rows with more than 3 contributions or more than 2 local
// copies (element corners on rank edges) take the CSR walk.
template <typename T, int kRP>
__global__ void __launch_bounds__(kGsBS, 2048 / kGsBS)
    gs_rows_halo_kernel(const GsRowsDev R, const GsGrid G, const int localBlocks, const HaloExchangeDev H,
                        const T* __restrict__ partial, T* __restrict__ v)
{
  pdl_trigger();
  if ((int)blockIdx.x < localBlocks) {
    gs_rows_body<T, kRP>(R, G, blockIdx.x, v);
    return;
  }
  const int row = (blockIdx.x - localBlocks) * kGsBS + threadIdx.x;
  int4 rf = make_int4(0, 0, 0, -1), rl = make_int4(-1, -1, 0, 0);
  if (row < H.nRows) {
    rf = H.recvFlat[row];
    rl = H.rowLocal[row];
  }
  pdl_wait();
  for (int i = threadIdx.x; i < H.nPeers * kFlagSlots; i += blockDim.x) {
    volatile unsigned long long* f = H.myFlags + (size_t)H.peerRank[i / kFlagSlots] * kFlagSlots + i % kFlagSlots;
    const long long t0 = clock64();
    while (*f < H.epoch) {
      if (clock64() - t0 > (1ll << 34)) {  // ~8 s: a peer died; flag it instead of hanging the box
        if (H.err) *H.err = 1;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  if (row >= H.nRows) return;
  const volatile T* w = (const volatile T*)H.myWindow;
  if (rf.w > 0 && rl.z <= 2) {
    const T own = partial[row];
    const T a = w[rf.x];
    const T b = (rf.w == 3) ? w[rf.y] : T(0);
    T tot;
    if (rf.w == 2) {
      tot = (rf.z == 0) ? own + a : a + own;
    } else {  // three contributions, own partial at position rf.z, the two remote ones in ascending rank order
      const T c0 = (rf.z == 0) ? own : a;
      const T c1 = (rf.z == 0) ? a : (rf.z == 1 ? own : b);
      const T c2 = (rf.z == 2) ? own : b;
      tot = (c0 + c1) + c2;
    }
    v[rl.x] = tot;
    if (rl.y >= 0) v[rl.y] = tot;
    return;
  }
  T tot = T(0);
  bool first = true;
  for (int c = H.recvStarts[row]; c < H.recvStarts[row + 1]; ++c) {
    const int p = H.recvPeer[c];
    const T val = (p < 0) ? partial[row] : w[(size_t)H.peerRecvOffset[p] + H.recvSlot[c]];
    tot = first ? val : tot + val;
    first = false;
  }
  for (int c = H.rowStarts[row]; c < H.rowStarts[row + 1]; ++c) v[H.rowIds[c]] = tot;
}

template <typename T>
int gs_rows_launch(const GsRowsDev& R, int Nfields, dlong stride, T* q, cudaStream_t stream)
{
  const long total = (long)R.nPairs + R.nQuads + R.nOcts + R.nGen + R.nMasked;
  if (total == 0 || Nfields == 0) return NRSB_OK;
  auto blocks = [](long n, long per) { return (int)((n + per - 1) / per); };
  GsGrid G;
  G.quadBlocks = blocks(R.nQuads, kGsBS);
  G.octBlocks = blocks(R.nOcts, kGsBS);
  G.genBlocks = blocks(R.nGen, kGsBS);
  G.maskBlocks = blocks(R.nMasked, (long)kGsBS * kGsMaskPerThread);
  // pair rows per thread: the smallest of 1..3 (fp32: 1..4) that keeps the whole grid within one wave (NRSB_GS_RPT forces it)
  static const int forced = getenv("NRSB_GS_RPT") ? atoi(getenv("NRSB_GS_RPT")) : 0;
  const long wave = (long)kNumSMs * (2048 / kGsBS);
  int rp = 1;
  const long others = (long)(G.quadBlocks + G.octBlocks + G.genBlocks + G.maskBlocks) * Nfields;
  const int rpMax = sizeof(T) == 8 ? 3 : 4;  // fp64: 4 pair rows per thread do not fit 32 registers
  while (rp < rpMax && (long)blocks(R.nPairs, (long)kGsBS * rp) * Nfields + others > wave) ++rp;
  if (forced >= 1 && forced <= rpMax) rp = forced;
  G.pairBlocks = blocks(R.nPairs, (long)kGsBS * rp);
  dim3 grid((unsigned)(G.pairBlocks + G.quadBlocks + G.octBlocks + G.genBlocks + G.maskBlocks), Nfields);
  auto kern = rp == 1 ? gs_rows_kernel<T, 1> : rp == 2 ? gs_rows_kernel<T, 2> : rp == 3 ? gs_rows_kernel<T, 3>
                                                                                       : gs_rows_kernel<T, 4>;
  NRSB_CUDA(launch_pdl_consumer(kern, grid, dim3(kGsBS), 0, stream, R, G, stride, q));
  return NRSB_OK;
}
// ---- the whole oogs::startFinish (oogs.cpp:823-837: packBuf, exchange, gatherScatterMany, unpackBuf) for one field and
// ogsAdd in ONE launch.  The reference needs pack kernel -> device sync -> MPI_Isend/Irecv/Waitall -> unpack kernel;
// the fence + epoch-flag version of this file needs two launches and a system-scope fence (measured: ~15 us per
// exchange more than the single-rank gather-scatter, 50 exchanges per V-cycle).  Here every halo value travels WITH
// its flag: an 8-byte NVLink store {fp32 value | 32-bit epoch} (fp64: two such words, low and high half), which is
// single-copy atomic, so neither a fence nor a separate flag nor a second launch is needed.  One thread per halo row:
//   partial sum of the local copies -> store to every sharer's window -> poll the own window for the sharers' words
//   -> add in ascending rank order (own partial at its own position: identical bits on every rank) -> local copies.
// The on-rank rows and the mask are the other block kinds of the same grid (halo blocks first: their stores leave
// as early as possible).  Windows are double-buffered by epoch parity (a rank cannot be more than one exchange
// ahead of a neighbour because it needs the neighbour's words of the current one).
template <typename T>
__device__ __forceinline__ void ll_store(void* win, const int slot, const T val, const unsigned epoch)
{
  volatile unsigned long long* w = (volatile unsigned long long*)win + 2 * (size_t)slot;
  if constexpr (sizeof(T) == 4) {
    w[0] = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(val);
  } else {
    const unsigned long long b = (unsigned long long)__double_as_longlong(val);
    w[0] = ((unsigned long long)epoch << 32) | (b & 0xffffffffull);
    w[1] = ((unsigned long long)epoch << 32) | (b >> 32);
  }
}
template <typename T>
__device__ __forceinline__ T ll_poll(const void* win, const int slot, const unsigned epoch, int* err)
{
  const volatile unsigned long long* w = (const volatile unsigned long long*)win + 2 * (size_t)slot;
  const long long t0 = clock64();
  unsigned long long a = w[0], b = 0;
  if constexpr (sizeof(T) == 8) b = w[1];
  while ((unsigned)(a >> 32) != epoch || (sizeof(T) == 8 && (unsigned)(b >> 32) != epoch)) {
    if (clock64() - t0 > (1ll << 34)) {  // ~8 s: the peer never sent; flag it instead of hanging the box
      if (err) *err = 1;
      break;
    }
    a = w[0];
    if constexpr (sizeof(T) == 8) b = w[1];
  }
  if constexpr (sizeof(T) == 4)
    return __uint_as_float((unsigned)a);
  else
    return __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
}

template <typename T, int kRP>
__global__ void __launch_bounds__(kGsBS, 2048 / kGsBS)
    gs_exchange_ll_kernel(const GsRowsDev R, const GsGrid G, const int haloBlocks, const HaloExchangeDev H,
                          T* __restrict__ v)
{
  pdl_trigger();
  if ((int)blockIdx.x >= haloBlocks) {
    gs_rows_body<T, kRP>(R, G, blockIdx.x - haloBlocks, v);
    return;
  }
  const int row = blockIdx.x * kGsBS + threadIdx.x;
  if (row >= H.nRows) return;
  const int4 rl = H.rowLocal[row];  // {id0, id1 | -1, local copies, -}
  const int4 rs = H.rowSend[row];   // {peer, slot, destinations, -}
  const int4 rf = H.recvFlat[row];  // {slot a, slot b, own position, contributions | -1}
  pdl_wait();
  // own partial: local copies in ascending local index
  T own;
  if (rl.z <= 2) {
    own = v[rl.x];
    if (rl.y >= 0) own += v[rl.y];
  } else {
    const int s0 = H.rowStarts[row], s1 = H.rowStarts[row + 1];
    own = v[H.rowIds[s0]];
    for (int c = s0 + 1; c < s1; ++c) own += v[H.rowIds[c]];
  }
  // push to every sharer
  if (rs.z == 1) {
    ll_store<T>(H.peerWindowInline[rs.x], rs.y, own, H.epoch32);
  } else {
    for (int d = H.sendStarts[row]; d < H.sendStarts[row + 1]; ++d) {
      const int p = H.sendPeer[d];
      ll_store<T>(H.peerWindowInline[p], (int)H.peerRemoteOffset[p] + H.sendSlot[d], own, H.epoch32);
    }
  }
  // receive and fold in ascending rank order
  T tot;
  if (rf.w == 2) {
    const T a = ll_poll<T>(H.myLL, rf.x, H.epoch32, H.err);
    tot = (rf.z == 0) ? own + a : a + own;
  } else if (rf.w == 3) {
    const T a = ll_poll<T>(H.myLL, rf.x, H.epoch32, H.err);
    const T b = ll_poll<T>(H.myLL, rf.y, H.epoch32, H.err);
    const T c0 = (rf.z == 0) ? own : a;
    const T c1 = (rf.z == 0) ? a : (rf.z == 1 ? own : b);
    const T c2 = (rf.z == 2) ? own : b;
    tot = (c0 + c1) + c2;
  } else {
    tot = T(0);
    bool first = true;
    for (int c = H.recvStarts[row]; c < H.recvStarts[row + 1]; ++c) {
      const int p = H.recvPeer[c];
      const T val = (p < 0) ? own : ll_poll<T>(H.myLL, (int)H.peerRecvOffset[p] + H.recvSlot[c], H.epoch32, H.err);
      tot = first ? val : tot + val;
      first = false;
    }
  }
  if (rl.z <= 2) {
    v[rl.x] = tot;
    if (rl.y >= 0) v[rl.y] = tot;
  } else {
    for (int c = H.rowStarts[row]; c < H.rowStarts[row + 1]; ++c) v[H.rowIds[c]] = tot;
  }
}

template <typename T>
int gs_exchange_ll_launch(const GsRowsDev& R, const HaloExchangeDev& H, T* v, cudaStream_t stream)
{
  auto blocks = [](long n, long per) { return (int)((n + per - 1) / per); };
  GsGrid G;
  G.quadBlocks = blocks(R.nQuads, kGsBS);
  G.octBlocks = blocks(R.nOcts, kGsBS);
  G.genBlocks = blocks(R.nGen, kGsBS);
  G.maskBlocks = blocks(R.nMasked, (long)kGsBS * kGsMaskPerThread);
  const int haloBlocks = blocks(H.nRows, kGsBS);
  const long wave = (long)kNumSMs * (2048 / kGsBS);
  const int rpMax = sizeof(T) == 8 ? 3 : 4;
  int rp = 1;
  const long others = (long)G.quadBlocks + G.octBlocks + G.genBlocks + G.maskBlocks + haloBlocks;
  while (rp < rpMax && (long)blocks(R.nPairs, (long)kGsBS * rp) + others > wave) ++rp;
  G.pairBlocks = blocks(R.nPairs, (long)kGsBS * rp);
  const int localBlocks = G.pairBlocks + G.quadBlocks + G.octBlocks + G.genBlocks + G.maskBlocks;
  auto kern = rp == 1 ? gs_exchange_ll_kernel<T, 1> : rp == 2 ? gs_exchange_ll_kernel<T, 2>
                      : rp == 3 ? gs_exchange_ll_kernel<T, 3> : gs_exchange_ll_kernel<T, 4>;
  NRSB_CUDA(launch_pdl_consumer(kern, dim3(localBlocks + haloBlocks), dim3(kGsBS), 0, stream, R, G, haloBlocks, H, v));
  return NRSB_OK;
}
template int gs_exchange_ll_launch<double>(const GsRowsDev&, const HaloExchangeDev&, double*, cudaStream_t);
template int gs_exchange_ll_launch<float>(const GsRowsDev&, const HaloExchangeDev&, float*, cudaStream_t);

template <typename T>
int gs_rows_halo_launch(const GsRowsDev& R, const HaloExchangeDev& H, const T* partial, T* v, cudaStream_t stream)
{
  auto blocks = [](long n, long per) { return (int)((n + per - 1) / per); };
  GsGrid G;
  G.quadBlocks = blocks(R.nQuads, kGsBS);
  G.octBlocks = blocks(R.nOcts, kGsBS);
  G.genBlocks = blocks(R.nGen, kGsBS);
  G.maskBlocks = blocks(R.nMasked, (long)kGsBS * kGsMaskPerThread);
  const int haloBlocks = blocks(H.nRows, kGsBS);
  const long wave = (long)kNumSMs * (2048 / kGsBS);
  const int rpMax = sizeof(T) == 8 ? 3 : 4;
  int rp = 1;
  const long others = (long)G.quadBlocks + G.octBlocks + G.genBlocks + G.maskBlocks + haloBlocks;
  while (rp < rpMax && (long)blocks(R.nPairs, (long)kGsBS * rp) + others > wave) ++rp;
  G.pairBlocks = blocks(R.nPairs, (long)kGsBS * rp);
  const int localBlocks = G.pairBlocks + G.quadBlocks + G.octBlocks + G.genBlocks + G.maskBlocks;
  auto kern = rp == 1 ? gs_rows_halo_kernel<T, 1> : rp == 2 ? gs_rows_halo_kernel<T, 2>
                      : rp == 3 ? gs_rows_halo_kernel<T, 3> : gs_rows_halo_kernel<T, 4>;
  NRSB_CUDA(launch_pdl_consumer(kern, dim3(localBlocks + haloBlocks), dim3(kGsBS), 0, stream, R, G, localBlocks, H,
                                partial, v));
  return NRSB_OK;
}
template int gs_rows_halo_launch<double>(const GsRowsDev&, const HaloExchangeDev&, const double*, double*, cudaStream_t);
template int gs_rows_halo_launch<float>(const GsRowsDev&, const HaloExchangeDev&, const float*, float*, cudaStream_t);

template int gs_rows_launch<double>(const GsRowsDev&, int, dlong, double*, cudaStream_t);
template int gs_rows_launch<float>(const GsRowsDev&, int, dlong, float*, cudaStream_t);

// ---- plain CSR gather-scatter (signature parity with gatherScatterMany.okl; used by the
//      kernel-level C ABI and by the setup-time checks)
template <typename T>
__global__ void __launch_bounds__(kBlockSize)
    gs_csr_kernel(const dlong Ngather, const int Nentries, const dlong stride, const dlong* __restrict__ starts,
                  const dlong* __restrict__ ids, T* __restrict__ q)
{
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)Ngather * Nentries) return;
  const dlong gid = g % Ngather;
  const int k = g / Ngather;
  const dlong start = starts[gid], end = starts[gid + 1];
  if (start + 1 == end) return;
  T gq = T(0);
  for (dlong n = start; n < end; ++n) gq += q[ids[n] + (size_t)k * stride];
  for (dlong n = start; n < end; ++n) q[ids[n] + (size_t)k * stride] = gq;
}

template <typename T>
int gs_csr_launch(dlong Ngather, int Nentries, dlong stride, const dlong* starts, const dlong* ids, T* q,
                  cudaStream_t stream)
{
  const long total = (long)Ngather * Nentries;
  if (total == 0) return NRSB_OK;
  gs_csr_kernel<T><<<(unsigned)((total + kBlockSize - 1) / kBlockSize), kBlockSize, 0, stream>>>(
      Ngather, Nentries, stride, starts, ids, q);
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}
template int gs_csr_launch<double>(dlong, int, dlong, const dlong*, const dlong*, double*, cudaStream_t);
template int gs_csr_launch<float>(dlong, int, dlong, const dlong*, const dlong*, float*, cudaStream_t);

// ---- mask  (kernels/core/mask.okl)
template <typename T>
__global__ void __launch_bounds__(kBlockSize) mask_kernel(const dlong Nmasked, const dlong* __restrict__ maskIds,
                                                          T* __restrict__ q)
{
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < Nmasked) q[maskIds[n]] = T(0);
}
template <typename T>
int mask_launch(dlong Nmasked, const dlong* maskIds, T* q, cudaStream_t stream)
{
  if (Nmasked == 0) return NRSB_OK;
  mask_kernel<T><<<(Nmasked + kBlockSize - 1) / kBlockSize, kBlockSize, 0, stream>>>(Nmasked, maskIds, q);
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}
template int mask_launch<double>(dlong, const dlong*, double*, cudaStream_t);
template int mask_launch<float>(dlong, const dlong*, float*, cudaStream_t);

}  // namespace nrsb

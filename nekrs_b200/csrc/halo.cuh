// halo.cuh -- device view of one NVLink halo exchange (oogs.cu) and the per-row pack step, shared with
// the fused axhelm + halo-push kernel (axhelm_tma.cu).
#pragma once
#include "common.cuh"
#include "host.hpp"

namespace nrsb {

constexpr int kFlagSlots = 16;   // arrival flags per sender: pusher CTA pb of the fused launch raises slots pb, pb + nPush, ...
constexpr int kInlinePeers = 8;  // peer window pointers carried in the kernel parameters (no dependent load)

struct HaloExchangeDev {
  int nRows;
  const int* rowStarts;  // local copies CSR
  const int* rowIds;
  const int* sendStarts;  // per row: destinations
  const int* sendPeer;    // peer index
  const int* sendSlot;    // slot inside my block of that peer's window
  // flat send table for one field (k = 1): one entry per (row, destination), sorted by row.
  //   sendFlat[e] = {id0, id1 (-1: single local copy), peer | kSendFirst | kSendSlow, absolute slot in the peer window}
  //   (kSendSlow: more than two local copies, id0 = row index -> CSR walk);  sendRow[e] = row
  //   kSendQuad: three or four local copies, the third and fourth id in sendExtra[e] (element corners on a rank face)
  int nSend;
  const int4* sendFlat;
  const int2* sendExtra;
  const int* sendRow;
  // flat receive table for one field: recvFlat[row] = {absolute slot a, absolute slot b, position of the own
  // partial among the contributions (ascending rank), number of contributions}; valid when that number <= 3,
  // else -1 in .w -> CSR walk.  rowLocal[row] = {id0, id1 (-1), number of local copies, 0}
  const int4* recvFlat;
  const int4* rowLocal;
  const int* recvStarts;  // per row: contributions in ascending rank order
  const int* recvPeer;    // peer index or -1 for the own partial
  const int* recvSlot;
  int nPeers;
  const long* peerRemoteOffset;  // my block's offset (slots) in peer's window
  const long* peerRecvOffset;    // peer's block offset (slots) in my window
  const int* peerCount;          // shared rows with peer
  const int* peerRank;
  void* const* peerWindow;  // this parity
  void* peerWindowInline[kInlinePeers];
  void* myWindow;           // this parity
  unsigned long long* const* peerFlags;
  unsigned long long* myFlags;
  unsigned* ticket;
  int myRank;
  unsigned long long epoch;
  int* err;  // host-mapped word: set when a wait for a peer timed out (checked by the host at its next sync)
  // one-launch exchange (gs_exchange_ll_kernel): flag-in-data windows, 16 bytes per slot, this parity; the peers'
  // windows come in peerWindowInline (<= kInlinePeers peers)
  const int4* rowSend;  // per halo row {peer, absolute slot in the peer window, number of destinations, 0}
  void* const* peerLL;
  void* myLL;
  unsigned epoch32;
};

struct GsRowsDev;
template <typename T>
int gs_rows_halo_launch(const GsRowsDev& R, const HaloExchangeDev& H, const T* partial, T* v, cudaStream_t stream);
template <typename T>
int gs_exchange_ll_launch(const GsRowsDev& R, const HaloExchangeDev& H, T* v, cudaStream_t stream);

template <typename T>
__device__ __forceinline__ T gs_combine(T a, T b, gs_op op)
{
  return op == gs_op::add ? a + b : (op == gs_op::min ? (b < a ? b : a) : (b > a ? b : a));
}

// partial sum of the local copies of halo row `row` (field f), stored locally and pushed into every
// sharer's receive window (packBuf of okl/oogs.okl:1-120 + the MPI send it feeds).
// kL2: read v through L2 (the values may have been written by other SMs of the same launch).
template <typename T, bool kL2 = false>
__device__ __forceinline__ void halo_pack_row(const HaloExchangeDev& H, const int k, const dlong stride, const gs_op op,
                                              const T* __restrict__ v, T* __restrict__ partial, const int row,
                                              const int f)
{
  const int s0 = H.rowStarts[row], s1 = H.rowStarts[row + 1];
  auto ld = [&](int c) -> T {
    const T* a = v + H.rowIds[c] + (size_t)f * stride;
    return kL2 ? __ldcg(a) : *a;
  };
  T s = ld(s0);
  for (int c = s0 + 1; c < s1; ++c) s = gs_combine(s, ld(c), op);
  partial[(size_t)f * H.nRows + row] = s;
  for (int d = H.sendStarts[row]; d < H.sendStarts[row + 1]; ++d) {
    const int p = H.sendPeer[d];
    T* w = (T*)H.peerWindow[p];
    w[(size_t)H.peerRemoteOffset[p] * k + (size_t)f * H.peerCount[p] + H.sendSlot[d]] = s;  // NVLink store
  }
}

constexpr int kSendFirst = 1 << 16;  // first destination of its row: also stores partial[row]
constexpr int kSendSlow = 1 << 17;
constexpr int kSendQuad = 1 << 18;

// kB send entries per thread in two steps, so that a caller can issue the index loads BEFORE the data is
// ready (they do not depend on it) and keep only  value load -> NVLink store  on the critical path.
template <int kB>
struct HaloSendBatch {
  int4 s[kB];
  int2 x[kB];
  int row[kB];
};

template <int kB>
__device__ __forceinline__ void halo_pack_load(const HaloExchangeDev& H, const int e0, const int estride,
                                               HaloSendBatch<kB>& b)
{
#pragma unroll
  for (int j = 0; j < kB; ++j) {
    const int e = e0 + j * estride;
    b.s[j] = make_int4(-1, -1, 0, 0);
    b.x[j] = make_int2(-1, -1);
    b.row[j] = 0;
    if (e < H.nSend) {
      b.s[j] = H.sendFlat[e];
      b.x[j] = H.sendExtra[e];
      b.row[j] = H.sendRow[e];
    }
  }
}

template <typename T, int kB, bool kL2>
__device__ __forceinline__ void halo_pack_gather(const HaloExchangeDev& H, const gs_op op, const T* __restrict__ v,
                                                 const int e0, const int estride, const HaloSendBatch<kB>& b,
                                                 T (&val)[kB])
{
  auto ld = [&](int id) -> T { return kL2 ? __ldcg(v + id) : v[id]; };
  // ALL value loads of the batch first (up to 4 local copies per entry, predicated: no branch between them), so
  // that a thread pays one L2 round trip for its kB entries instead of one per entry (measured on the pusher CTAs
  // of the fused launch: 13 us for 16 entries per thread with the loads behind per-entry branches)
  T c[kB][4];
#pragma unroll
  for (int j = 0; j < kB; ++j) {
    const bool on = (e0 + j * estride < H.nSend) && !(b.s[j].z & kSendSlow);
    const bool quad = on && (b.s[j].z & kSendQuad);
    c[j][0] = on ? ld(b.s[j].x) : T(0);
    c[j][1] = (on && b.s[j].y >= 0) ? ld(b.s[j].y) : T(0);
    c[j][2] = quad ? ld(b.x[j].x) : T(0);
    c[j][3] = (quad && b.x[j].y >= 0) ? ld(b.x[j].y) : T(0);
  }
#pragma unroll
  for (int j = 0; j < kB; ++j) {
    val[j] = T(0);
    if (e0 + j * estride < H.nSend) {
      if (b.s[j].z & kSendSlow) {  // more than four local copies: CSR walk (not on a conforming hex mesh interior)
        const int s0 = H.rowStarts[b.row[j]], s1 = H.rowStarts[b.row[j] + 1];
        T a = ld(H.rowIds[s0]);
        for (int cc = s0 + 1; cc < s1; ++cc) a = gs_combine(a, ld(H.rowIds[cc]), op);
        val[j] = a;
      } else {
        T sum = c[j][0];
        if (b.s[j].y >= 0) sum = gs_combine(sum, c[j][1], op);
        if (b.s[j].z & kSendQuad) {
          sum = gs_combine(sum, c[j][2], op);
          if (b.x[j].y >= 0) sum = gs_combine(sum, c[j][3], op);
        }
        val[j] = sum;
      }
    }
  }
}

template <typename T, int kB>
__device__ __forceinline__ void halo_pack_scatter(const HaloExchangeDev& H, T* __restrict__ partial, const int e0,
                                                  const int estride, const HaloSendBatch<kB>& b, const T (&val)[kB])
{
#pragma unroll
  for (int j = 0; j < kB; ++j)
    if (e0 + j * estride < H.nSend) {
      if (b.s[j].z & kSendFirst) partial[b.row[j]] = val[j];
      const int p = b.s[j].z & 0xffff;
      T* w = (T*)(p < kInlinePeers ? H.peerWindowInline[p] : H.peerWindow[p]);
      w[b.s[j].w] = val[j];  // NVLink store
    }
}

template <typename T, int kB, bool kL2>
__device__ __forceinline__ void halo_pack_store(const HaloExchangeDev& H, const gs_op op, const T* __restrict__ v,
                                                T* __restrict__ partial, const int e0, const int estride,
                                                const HaloSendBatch<kB>& b)
{
  T val[kB];
  halo_pack_gather<T, kB, kL2>(H, op, v, e0, estride, b, val);
  halo_pack_scatter<T, kB>(H, partial, e0, estride, b, val);
}

template <typename T, int kB, bool kL2>
__device__ __forceinline__ void halo_pack_flat(const HaloExchangeDev& H, const gs_op op, const T* __restrict__ v,
                                               T* __restrict__ partial, const int e0, const int estride)
{
  HaloSendBatch<kB> b;
  halo_pack_load<kB>(H, e0, estride, b);
  halo_pack_store<T, kB, kL2>(H, op, v, partial, e0, estride, b);
}

// what the fused axhelm kernel needs besides the exchange itself
struct FusedHalo {
  HaloExchangeDev H;
  int nPush = 8;                       // CTAs of the launch that only push halo sums (set by oogs::begin_fused)
  dlong NhaloElements = 0;             // the first NhaloElements entries of the element list touch halo rows
  unsigned long long* counter = nullptr;  // monotonically increasing count of finished halo elements
  unsigned long long target = 0;          // value of *counter once this launch's halo elements are all stored
  void* partial = nullptr;
  dlong stride = 0;
  // developer aid (NRSB_OP_TIMING): globaltimer stamps, [0..4] pusher 0 (start, halo elements stored, pushed,
  // fenced, flags raised), [8] first / [9] last axhelm CTA done
  unsigned long long* stamps = nullptr;
  int* err = nullptr;  // host-mapped error word: a bounded wait of this launch gave up
};

}  // namespace nrsb

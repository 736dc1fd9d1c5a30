// halo.cuh -- device view of one NVLink halo exchange (oogs.cu) and the per-row pack step, shared with
// the fused axhelm + halo-push kernel (axhelm_tma.cu).
#pragma once
#include "common.cuh"
#include "host.hpp"

namespace nrsb {

struct HaloExchangeDev {
  int nRows;
  const int* rowStarts;  // local copies CSR
  const int* rowIds;
  const int* sendStarts;  // per row: destinations
  const int* sendPeer;    // peer index
  const int* sendSlot;    // slot inside my block of that peer's window
  const int* recvStarts;  // per row: contributions in ascending rank order
  const int* recvPeer;    // peer index or -1 for the own partial
  const int* recvSlot;
  int nPeers;
  const long* peerRemoteOffset;  // my block's offset (slots) in peer's window
  const long* peerRecvOffset;    // peer's block offset (slots) in my window
  const int* peerCount;          // shared rows with peer
  const int* peerRank;
  void* const* peerWindow;  // this parity
  void* myWindow;           // this parity
  unsigned long long* const* peerFlags;
  unsigned long long* myFlags;
  unsigned* ticket;
  int myRank;
  unsigned long long epoch;
};

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

// what the fused axhelm kernel needs besides the exchange itself
struct FusedHalo {
  HaloExchangeDev H;
  dlong NhaloElements = 0;             // the first NhaloElements entries of the element list touch halo rows
  unsigned long long* counter = nullptr;  // monotonically increasing count of finished halo elements
  unsigned long long target = 0;          // value of *counter once this launch's halo elements are all stored
  void* partial = nullptr;
  dlong stride = 0;
};

}  // namespace nrsb

// oogs.cu -- overlapped gather-scatter with a device-initiated NVLink halo exchange.
//
// Replaces oogs::setup/start/finish (3rd_party/gslib/ogs/src/oogs.cpp:336-837) and the device side
// of its exchange, packBuf_*/unpackBuf_* (okl/oogs.okl:1-272).  The reference packs halo partial
// sums into a send buffer, synchronises the device, moves the buffer with MPI_Isend/Irecv (through
// pinned host memory unless GPU-aware MPI is enabled) and unpacks.  On an NVSwitch box every peer
// is one hop away, so here:
//
//   start : ONE kernel sums the local copies of every halo row and STORES the partial sum directly
//           into each sharer's receive window (peer-mapped memory opened once via CUDA IPC); the
//           last block to finish publishes an epoch flag to every peer.  No host involvement.
//   finish: ONE kernel waits for the peers' flags, adds the received partials in ascending rank
//           order (own contribution at its own position, so every rank computes bit-identical
//           sums), scatters the totals to the local copies, and in the same launch performs the
//           on-rank gather-scatter rows and the Dirichlet mask.
//
// Work launched on the same stream between start and finish (interior-element Ax) overlaps the
// NVLink traffic exactly like the reference's oogs::start / callback / oogs::finish sequence.
// Receive windows are double-buffered by epoch parity; a rank cannot be more than one exchange
// ahead of a neighbour because finish(e) needs the neighbour's flag e.
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <map>

#include "halo.cuh"
#include "host.hpp"

namespace nrsb {

template <typename T>
__global__ void __launch_bounds__(kBlockSize)
    halo_pack_kernel(const HaloExchangeDev H, const int k, const dlong stride, const gs_op op, const T* __restrict__ v,
                     T* __restrict__ partial)
{
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 1) {
    halo_pack_flat<T, 1, false>(H, op, v, partial, (int)gid, 0);
  } else if (gid < (long)H.nRows * k) {
    halo_pack_row<T>(H, k, stride, op, v, partial, (int)(gid % H.nRows), (int)(gid / H.nRows));
  }
  // publish: every block fences its remote stores, the last one raises the flags
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(H.ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence_system();
    for (int p = threadIdx.x; p < H.nPeers * kFlagSlots; p += blockDim.x) {
      volatile unsigned long long* f = H.peerFlags[p / kFlagSlots] + (size_t)H.myRank * kFlagSlots + p % kFlagSlots;
      *f = H.epoch;
    }
    if (threadIdx.x == 0) *H.ticket = 0u;
  }
}

template <typename T>
__device__ __forceinline__ void unpack_rows_body(const HaloExchangeDev& H, const GsRowsDev& R, const int k,
                                                 const dlong stride, const gs_op op, const T* __restrict__ partial,
                                                 T* __restrict__ v, const int haloBlocks)
{
  if ((int)blockIdx.x >= haloBlocks) {
    // ---- on-rank rows + mask (same code path as gs_rows_kernel, add only)
    const int n = (blockIdx.x - haloBlocks) * blockDim.x + threadIdx.x;
    T* __restrict__ qf = v + (size_t)blockIdx.y * stride;
    if (n < R.nPairs) {
      const int2 id = R.pairs[n];
      const T s = gs_combine(qf[id.x], qf[id.y], op);
      qf[id.x] = s;
      qf[id.y] = s;
      return;
    }
    int m = n - R.nPairs;
    if (m < R.nQuads) {
      const int4 id = R.quads[m];
      T s = gs_combine(qf[id.x], qf[id.y], op);
      s = gs_combine(s, qf[id.z], op);
      s = gs_combine(s, qf[id.w], op);
      qf[id.x] = s;
      qf[id.y] = s;
      qf[id.z] = s;
      qf[id.w] = s;
      return;
    }
    m -= R.nQuads;
    if (m < R.nOcts) {
      const int4 ia = R.octs[2 * m], ib = R.octs[2 * m + 1];
      T s = gs_combine(qf[ia.x], qf[ia.y], op);
      s = gs_combine(s, qf[ia.z], op);
      s = gs_combine(s, qf[ia.w], op);
      s = gs_combine(s, qf[ib.x], op);
      s = gs_combine(s, qf[ib.y], op);
      s = gs_combine(s, qf[ib.z], op);
      s = gs_combine(s, qf[ib.w], op);
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
    m -= R.nOcts;
    if (m < R.nGen) {
      const int start = R.genStarts[m], end = R.genStarts[m + 1];
      T s = qf[R.genIds[start]];
      for (int c = start + 1; c < end; ++c) s = gs_combine(s, qf[R.genIds[c]], op);
      for (int c = start; c < end; ++c) qf[R.genIds[c]] = s;
      return;
    }
    m -= R.nGen;
    if (m < R.nMasked) qf[R.maskIds[m]] = T(0);
    return;
  }
  // ---- halo rows: wait for every peer's epoch flag, then fold
  if (blockIdx.y != 0) return;  // halo blocks handle all fields themselves
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
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long)H.nRows * k) return;
  const int row = gid % H.nRows;
  const int f = gid / H.nRows;
  const volatile T* w = (const volatile T*)H.myWindow;
  if (k == 1) {
    // common case through the flat tables: two independent index loads, then the values, then the stores
    const int4 rf = H.recvFlat[row];
    const int4 rl = H.rowLocal[row];
    if (rf.w > 0 && rl.z <= 2) {
      const T own = partial[row];
      const T a = w[rf.x];
      const T b = (rf.w == 3) ? w[rf.y] : T(0);
      T tot;
      if (rf.w == 2) {
        tot = (rf.z == 0) ? gs_combine(own, a, op) : gs_combine(a, own, op);
      } else {  // three contributions, own partial at position rf.z, the two remote ones in ascending rank order
        const T c0 = (rf.z == 0) ? own : a;
        const T c1 = (rf.z == 0) ? a : (rf.z == 1 ? own : b);
        const T c2 = (rf.z == 2) ? own : b;
        tot = gs_combine(gs_combine(c0, c1, op), c2, op);
      }
      v[rl.x] = tot;
      if (rl.y >= 0) v[rl.y] = tot;
      return;
    }
  }
  T tot = T(0);
  bool first = true;
  for (int c = H.recvStarts[row]; c < H.recvStarts[row + 1]; ++c) {
    const int p = H.recvPeer[c];
    const T val = (p < 0) ? partial[(size_t)f * H.nRows + row]
                          : w[(size_t)H.peerRecvOffset[p] * k + (size_t)f * H.peerCount[p] + H.recvSlot[c]];
    tot = first ? val : gs_combine(tot, val, op);
    first = false;
  }
  for (int c = H.rowStarts[row]; c < H.rowStarts[row + 1]; ++c) v[H.rowIds[c] + (size_t)f * stride] = tot;
}

template <typename T>
__global__ void __launch_bounds__(kBlockSize)
    halo_unpack_rows_kernel(const HaloExchangeDev H, const GsRowsDev R, const int k, const dlong stride,
                            const gs_op op, const T* __restrict__ partial, T* __restrict__ v, const int haloBlocks)
{
  unpack_rows_body<T>(H, R, k, stride, op, partial, v, haloBlocks);
}

// ---- on-rank only, any op (add uses the faster gs_rows_kernel in gs.cu)
template <typename T>
__global__ void __launch_bounds__(kBlockSize)
    rows_op_kernel(const HaloExchangeDev H, const GsRowsDev R, const int k, const dlong stride, const gs_op op,
                   T* __restrict__ v)
{
  unpack_rows_body<T>(H, R, k, stride, op, nullptr, v, 0);
}

// ------------------------------------------------------------------------------------------
struct oogs_dev_t {
  std::vector<void*> h_peerWindow[2];
  dbuf<int4> sendFlat, recvFlat, rowLocal;
  dbuf<int2> sendExtra;
  dbuf<int> sendRow;
  dbuf<long> peerRemoteOffset, peerRecvOffset;
  dbuf<int> peerCount, peerRank;
  dbuf<int> recvPeer, recvSlot;
  dbuf<unsigned> ticket;
  dbuf<void*> d_peerWindow[2];
  void* arena = nullptr;
  std::vector<void*> peerArena;  // IPC-opened
  size_t windowBytes = 0;
  // flag-in-data windows of the one-launch exchange (gs.cu gs_exchange_ll_kernel): 16 bytes per slot, two parities,
  // behind the flags in the same arena (one IPC handle)
  size_t llOffset = 0, llBytes = 0;
  std::vector<void*> h_peerLL[2];
  dbuf<void*> d_peerLL[2];
  dbuf<int4> rowSend;  // per halo row {peer, absolute slot in the peer window, number of destinations, 0}
  unsigned long long epochLL = 0;
};
static std::map<oogs_t*, std::unique_ptr<oogs_dev_t>> g_dev;

oogs_t::~oogs_t()
{
  auto it = g_dev.find(this);
  if (it != g_dev.end()) {
    for (size_t p = 0; p < it->second->peerArena.size(); ++p)
      if (it->second->peerArena[p]) cudaIpcCloseMemHandle(it->second->peerArena[p]);
    cudaFree(it->second->arena);
    g_dev.erase(it);
  }
}

int oogs_t::setup(ogs_t* ogs_, comm_t* comm_, int maxFields_)
{
  ogs = ogs_;
  comm = comm_;
  maxFields = std::max(1, maxFields_);
  if (!comm || comm->nranks == 1) {
    if (ogs->NhaloGather) {
      set_last_error("ogs has halo rows but no communicator was given");
      return NRSB_ERR_INVALID;
    }
    return NRSB_OK;
  }
  if (ogs->NhaloGather == 0) {
    // A rank without halo rows in THIS handle (all its interface nodes Dirichlet-masked, a coarse level that lives on
    // fewer ranks, ...) still has to take part in the setup collectives of the ranks that have some: same sequence
    // as below with empty contributions (counts, a dummy IPC handle nobody opens, the two barriers).
    const int nr = comm->nranks;
    std::vector<int> C0((size_t)nr * nr, 0);
    comm->allgather_bytes(C0.data(), sizeof(int) * nr);
    void* dummy = nullptr;
    NRSB_CUDA(cudaMalloc(&dummy, 256));
    std::vector<cudaIpcMemHandle_t> h0(nr);
    NRSB_CUDA(cudaIpcGetMemHandle(&h0[comm->rank], dummy));
    comm->allgather_bytes(h0.data(), sizeof(cudaIpcMemHandle_t));
    comm->barrier();
    comm->barrier();
    cudaFree(dummy);
    return NRSB_OK;
  }
  const int nranks = comm->nranks, rank = comm->rank;
  const int nRows = ogs->NhaloGather;
  int rc;

  // rows shared with each peer in ascending global id order
  std::vector<int> order(nRows);
  for (int r = 0; r < nRows; ++r) order[r] = r;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return ogs->haloBaseIds[a] < ogs->haloBaseIds[b]; });
  std::vector<std::vector<int>> rowsOf(nranks);
  for (int r : order)
    for (int c = ogs->haloSharerOffsets[r]; c < ogs->haloSharerOffsets[r + 1]; ++c) {
      const int q = ogs->haloSharerRanks[c];
      if (q != rank) rowsOf[q].push_back(r);
    }
  // counts matrix C[r][q]
  std::vector<int> C((size_t)nranks * nranks, 0);
  for (int q = 0; q < nranks; ++q) C[(size_t)rank * nranks + q] = (int)rowsOf[q].size();
  comm->allgather_bytes(C.data(), sizeof(int) * nranks);
  for (int q = 0; q < nranks; ++q)
    if (C[(size_t)rank * nranks + q] != C[(size_t)q * nranks + rank]) {
      set_last_error("halo topology is not symmetric between ranks");
      return NRSB_ERR_INVALID;
    }
  peers.clear();
  std::vector<int> peerIndexOfRank(nranks, -1);
  size_t off = 0;
  for (int q = 0; q < nranks; ++q) {
    if (q == rank || rowsOf[q].empty()) continue;
    Peer p;
    p.rank = q;
    p.nSend = (int)rowsOf[q].size();
    p.rows = rowsOf[q];
    p.recvOffset = off;
    off += p.nSend;
    // my block's offset inside q's window: q orders its peers by ascending rank too
    size_t ro = 0;
    for (int qq = 0; qq < rank; ++qq)
      if (qq != q) ro += C[(size_t)q * nranks + qq];
    p.remoteOffset = ro;
    peerIndexOfRank[q] = (int)peers.size();
    peers.push_back(p);
  }
  windowSlots = off;

  auto dev = std::make_unique<oogs_dev_t>();
  // arena = [window parity 0][window parity 1][flags nranks]
  dev->windowBytes = ((windowSlots * maxFields * sizeof(double) + 255) / 256) * 256;
  dev->llOffset = ((2 * dev->windowBytes + sizeof(unsigned long long) * nranks * kFlagSlots + 255) / 256) * 256;
  dev->llBytes = ((windowSlots * 16 + 255) / 256) * 256;
  const size_t arenaBytes = dev->llOffset + 2 * dev->llBytes;
  NRSB_CUDA(cudaMalloc(&dev->arena, arenaBytes));
  NRSB_CUDA(cudaMemset(dev->arena, 0, arenaBytes));
  NRSB_CUDA(cudaDeviceSynchronize());
  // exchange IPC handles
  std::vector<cudaIpcMemHandle_t> handles(nranks);
  NRSB_CUDA(cudaIpcGetMemHandle(&handles[rank], dev->arena));
  comm->allgather_bytes(handles.data(), sizeof(cudaIpcMemHandle_t));
  dev->peerArena.assign(nranks, nullptr);
  for (auto& p : peers)
    NRSB_CUDA(cudaIpcOpenMemHandle(&dev->peerArena[p.rank], handles[p.rank], cudaIpcMemLazyEnablePeerAccess));
  comm->barrier();

  // device tables
  std::vector<long> remOff, recOff;
  std::vector<int> cnt, prank;
  std::vector<void*> win[2];
  std::vector<unsigned long long*> pflags;
  for (auto& p : peers) {
    remOff.push_back((long)p.remoteOffset);
    recOff.push_back((long)p.recvOffset);
    cnt.push_back(p.nSend);
    prank.push_back(p.rank);
    char* base = (char*)dev->peerArena[p.rank];
    win[0].push_back(base);
    win[1].push_back(base + dev->windowBytes);
    pflags.push_back((unsigned long long*)(base + 2 * dev->windowBytes));
  }
  if ((rc = dev->peerRemoteOffset.upload(remOff))) return rc;
  if ((rc = dev->peerRecvOffset.upload(recOff))) return rc;
  if ((rc = dev->peerCount.upload(cnt))) return rc;
  if ((rc = dev->peerRank.upload(prank))) return rc;
  dev->h_peerWindow[0] = win[0];
  dev->h_peerWindow[1] = win[1];
  for (int par = 0; par < 2; ++par) {
    for (auto& p : peers)
      dev->h_peerLL[par].push_back((char*)dev->peerArena[p.rank] + dev->llOffset + (size_t)par * dev->llBytes);
    if ((rc = dev->d_peerLL[par].upload(dev->h_peerLL[par]))) return rc;
  }
  if ((rc = dev->d_peerWindow[0].upload(win[0]))) return rc;
  if ((rc = dev->d_peerWindow[1].upload(win[1]))) return rc;
  if ((rc = d_peerFlags.upload(pflags))) return rc;
  if ((rc = dev->ticket.alloc(1))) return rc;

  // send / receive CSR per row
  std::vector<int> sendStarts(nRows + 1, 0), sendPeer, sendSlot, recvStarts(nRows + 1, 0), recvPeer, recvSlot;
  std::vector<std::map<int, int>> slotOf(peers.size());  // row -> slot per peer
  for (size_t pi = 0; pi < peers.size(); ++pi)
    for (int s = 0; s < peers[pi].nSend; ++s) slotOf[pi][peers[pi].rows[s]] = s;
  for (int r = 0; r < nRows; ++r) {
    for (int c = ogs->haloSharerOffsets[r]; c < ogs->haloSharerOffsets[r + 1]; ++c) {
      const int q = ogs->haloSharerRanks[c];
      if (q == rank) {
        recvPeer.push_back(-1);
        recvSlot.push_back(0);
      } else {
        const int pi = peerIndexOfRank[q];
        const int slot = slotOf[pi][r];
        sendPeer.push_back(pi);
        sendSlot.push_back(slot);
        recvPeer.push_back(pi);
        recvSlot.push_back(slot);
      }
    }
    sendStarts[r + 1] = (int)sendPeer.size();
    recvStarts[r + 1] = (int)recvPeer.size();
  }
  {
    std::vector<int4> flat(sendPeer.size());
    std::vector<int2> fextra(sendPeer.size(), make_int2(-1, -1));
    std::vector<int> frow(sendPeer.size());
    for (int r = 0; r < nRows; ++r) {
      const int c0 = ogs->haloGatherOffsets[r], c1 = ogs->haloGatherOffsets[r + 1];
      for (int d = sendStarts[r]; d < sendStarts[r + 1]; ++d) {
        int code = sendPeer[d];
        if (d == sendStarts[r]) code |= kSendFirst;
        int id0 = ogs->haloGatherIds[c0], id1 = -1;
        if (c1 - c0 == 2) id1 = ogs->haloGatherIds[c0 + 1];
        if (c1 - c0 == 3 || c1 - c0 == 4) {
          code |= kSendQuad;
          id1 = ogs->haloGatherIds[c0 + 1];
          fextra[d] = make_int2(ogs->haloGatherIds[c0 + 2], c1 - c0 == 4 ? ogs->haloGatherIds[c0 + 3] : -1);
        } else if (c1 - c0 > 4) {
          code |= kSendSlow;
          id0 = r;
        }
        flat[d] = make_int4(id0, id1, code, (int)(peers[sendPeer[d]].remoteOffset + sendSlot[d]));
        frow[d] = r;
      }
    }
    if ((rc = dev->sendFlat.upload(flat))) return rc;
    if ((rc = dev->sendExtra.upload(fextra))) return rc;
    if ((rc = dev->sendRow.upload(frow))) return rc;
    std::vector<int4> rflat(nRows), rloc(nRows);
    for (int r = 0; r < nRows; ++r) {
      const int c0 = ogs->haloGatherOffsets[r], c1 = ogs->haloGatherOffsets[r + 1];
      rloc[r] = make_int4(ogs->haloGatherIds[c0], c1 - c0 >= 2 ? ogs->haloGatherIds[c0 + 1] : -1, c1 - c0, 0);
      const int n = recvStarts[r + 1] - recvStarts[r];
      int4 e = make_int4(0, 0, 0, -1);
      if (n == 2 || n == 3) {
        int slots[2] = {0, 0}, ns = 0, ownPos = 0;
        for (int c = recvStarts[r]; c < recvStarts[r + 1]; ++c) {
          if (recvPeer[c] < 0)
            ownPos = c - recvStarts[r];
          else
            slots[ns++] = (int)(peers[recvPeer[c]].recvOffset + recvSlot[c]);
        }
        e = make_int4(slots[0], slots[1], ownPos, n);
      }
      rflat[r] = e;
    }
    if ((rc = dev->recvFlat.upload(rflat))) return rc;
    if ((rc = dev->rowLocal.upload(rloc))) return rc;
    std::vector<int4> rsend(nRows);
    for (int r = 0; r < nRows; ++r) {
      const int d0 = sendStarts[r], nd = sendStarts[r + 1] - d0;
      rsend[r] = make_int4(nd > 0 ? sendPeer[d0] : 0, nd > 0 ? (int)(peers[sendPeer[d0]].remoteOffset + sendSlot[d0]) : 0,
                           nd, 0);
    }
    if ((rc = dev->rowSend.upload(rsend))) return rc;
  }
  if ((rc = d_sendStarts.upload(sendStarts))) return rc;
  if ((rc = d_sendPeer.upload(sendPeer))) return rc;
  if ((rc = d_sendSlot.upload(sendSlot))) return rc;
  if ((rc = d_recvStarts.upload(recvStarts))) return rc;
  if ((rc = dev->recvPeer.upload(recvPeer))) return rc;
  if ((rc = dev->recvSlot.upload(recvSlot))) return rc;
  if ((rc = d_partial.alloc((size_t)nRows * maxFields))) return rc;
  epoch = 0;
  g_dev[this] = std::move(dev);

  // global multiplicities -> invDegree (ogsSetup.cpp:366-392): gs(add) of ones through this exchange
  {
    dbuf<double> ones;
    std::vector<double> h(ogs->N, 1.0);
    if ((rc = ones.upload(h))) return rc;
    if ((rc = startFinish<double>(ones.p, 1, 0, gs_op::add, 0, nullptr, nullptr))) return rc;
    NRSB_CUDA(cudaDeviceSynchronize());
    if ((rc = ones.download(h))) return rc;
    for (dlong n = 0; n < ogs->N; ++n) ogs->invDegree[n] = 1.0 / h[n];
    if ((rc = ogs->upload_inv_degree())) return rc;
  }
  comm->barrier();
  return NRSB_OK;
}

static HaloExchangeDev make_dev(oogs_t* o, oogs_dev_t* d, int parity)
{
  HaloExchangeDev H;
  H.nRows = o->ogs->NhaloGather;
  H.rowStarts = o->ogs->d_haloStarts;
  H.rowIds = o->ogs->d_haloIds;
  H.sendStarts = o->d_sendStarts.p;
  H.sendPeer = o->d_sendPeer.p;
  H.sendSlot = o->d_sendSlot.p;
  H.nSend = (int)d->sendFlat.n;
  H.sendFlat = d->sendFlat.p;
  H.sendExtra = d->sendExtra.p;
  H.sendRow = d->sendRow.p;
  H.recvFlat = d->recvFlat.p;
  H.rowLocal = d->rowLocal.p;
  H.recvStarts = o->d_recvStarts.p;
  H.recvPeer = d->recvPeer.p;
  H.recvSlot = d->recvSlot.p;
  H.nPeers = (int)o->peers.size();
  H.peerRemoteOffset = d->peerRemoteOffset.p;
  H.peerRecvOffset = d->peerRecvOffset.p;
  H.peerCount = d->peerCount.p;
  H.peerRank = d->peerRank.p;
  H.peerWindow = d->d_peerWindow[parity].p;
  for (int p = 0; p < kInlinePeers; ++p)
    H.peerWindowInline[p] = p < (int)d->h_peerWindow[parity].size() ? d->h_peerWindow[parity][p] : nullptr;
  H.myWindow = (char*)d->arena + (size_t)parity * d->windowBytes;
  H.peerFlags = o->d_peerFlags.p;
  H.myFlags = (unsigned long long*)((char*)d->arena + 2 * d->windowBytes);
  H.ticket = d->ticket.p;
  H.myRank = o->comm->rank;
  H.epoch = o->epoch;
  H.err = o->comm ? o->comm->d_err : nullptr;
  H.rowSend = d->rowSend.p;
  H.peerLL = nullptr;
  H.myLL = nullptr;
  H.epoch32 = 0;
  return H;
}

template <typename T>
int oogs_t::start(T* v, int k, dlong stride, gs_op op, cudaStream_t stream)
{
  if (!ogs || ogs->NhaloGather == 0) return NRSB_OK;
  NRSB_REQUIRE(k <= maxFields, "oogs::start: more fields than the handle was set up for");
  NRSB_REQUIRE((int)peers.size() <= 64, "too many neighbour ranks");
  oogs_dev_t* d = g_dev[this].get();
  ++epoch;
  HaloExchangeDev H = make_dev(this, d, (int)(epoch & 1ull));
  const long total = (k == 1) ? (long)H.nSend : (long)H.nRows * k;
  halo_pack_kernel<T><<<(unsigned)((total + kBlockSize - 1) / kBlockSize), kBlockSize, 0, stream>>>(
      H, k, stride, op, v, (T*)d_partial.p);
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

int oogs_t::begin_fused(FusedHalo* F, dlong NhaloElements, dlong stride)
{
  NRSB_REQUIRE(ogs && ogs->NhaloGather > 0, "begin_fused: no halo rows");
  NRSB_REQUIRE((int)peers.size() <= 32, "begin_fused: too many neighbour ranks");
  oogs_dev_t* d = g_dev[this].get();
  if (!fusedCounter.p) {
    int rc = fusedCounter.alloc(1);
    if (rc) return rc;
    NRSB_CUDA(cudaMemset(fusedCounter.p, 0, sizeof(unsigned long long)));
  }
  ++epoch;
  fusedTarget += (unsigned long long)NhaloElements;
  F->H = make_dev(this, d, (int)(epoch & 1ull));
  // about 8 send entries per pusher thread (224 threads per CTA), between 2 and kFlagSlots pushers; every pusher
  // is an SM the element work does not get.  Measured at 2 GPUs (12.8 K entries): the value loads of a pusher CTA
  // take 10 us with 16 entries per thread and 5 us with 8, so that the flags go up at 16 us instead of 20 us after
  // the launch started -- before the element work ends (19-20 us) instead of after it.
  {
    static const int forced = getenv("NRSB_NPUSH") ? atoi(getenv("NRSB_NPUSH")) : 0;
    int np = (int)((F->H.nSend + 224 * 8 - 1) / (224 * 8));
    np = np < 2 ? 2 : (np > kFlagSlots ? kFlagSlots : np);
    if (forced >= 1 && forced <= kFlagSlots) np = forced;
    F->nPush = np;
  }
  F->NhaloElements = NhaloElements;
  F->counter = fusedCounter.p;
  F->target = fusedTarget;
  F->partial = d_partial.p;
  F->stride = stride;
  return NRSB_OK;
}

template <typename T>
int oogs_t::finish(T* v, int k, dlong stride, gs_op op, dlong Nmasked, const dlong* maskIds, cudaStream_t stream)
{
  GsRowsDev R = ogs->rows;
  R.nMasked = maskIds ? Nmasked : 0;
  R.maskIds = maskIds;
  const long localWork = (long)R.nPairs + R.nQuads + R.nOcts + R.nGen + R.nMasked;
  if (ogs->NhaloGather == 0) {
    if (op == gs_op::add) return gs_rows_launch<T>(R, k, stride, v, stream);
    if (localWork == 0) return NRSB_OK;
    HaloExchangeDev H{};
    dim3 grid((unsigned)((localWork + kBlockSize - 1) / kBlockSize), k);
    rows_op_kernel<T><<<grid, kBlockSize, 0, stream>>>(H, R, k, stride, op, v);
    NRSB_CHECK_LAUNCH();
    return NRSB_OK;
  }
  oogs_dev_t* d = g_dev[this].get();
  HaloExchangeDev H = make_dev(this, d, (int)(epoch & 1ull));
  // the operator's case (one field, add): single-wave kernel with one row kind per block (gs.cu)
  if (k == 1 && op == gs_op::add && getenv("NRSB_OLD_FINISH") == nullptr)
    return gs_rows_halo_launch<T>(R, H, (const T*)d_partial.p, v, stream);
  const long haloWork = (long)H.nRows * k;
  const int haloBlocks = (int)((haloWork + kBlockSize - 1) / kBlockSize);
  const int localBlocks = (int)((localWork + kBlockSize - 1) / kBlockSize);
  dim3 grid(haloBlocks + localBlocks, k);
  halo_unpack_rows_kernel<T><<<grid, kBlockSize, 0, stream>>>(H, R, k, stride, op, (const T*)d_partial.p, v, haloBlocks);
  NRSB_CHECK_LAUNCH();
  return NRSB_OK;
}

// oogs::startFinish for one field and ogsAdd in ONE launch with flag-in-data windows: see gs_exchange_ll_kernel
// (gs.cu).  Returns 1 when the case is not covered (the caller then runs start + finish).
template <typename T>
int oogs_t::exchange_ll(T* v, int k, gs_op op, dlong Nmasked, const dlong* maskIds, cudaStream_t stream)
{
  static const bool off = getenv("NRSB_NO_LL_EXCHANGE") != nullptr;
  if (off || !ogs || ogs->NhaloGather == 0 || k != 1 || op != gs_op::add || peers.size() > (size_t)kInlinePeers) return 1;
  oogs_dev_t* d = g_dev[this].get();
  ++d->epochLL;
  if ((d->epochLL & 0xffffffffull) == 0) ++d->epochLL;  // 0 is "never written"
  const int parity = (int)(d->epochLL & 1ull);
  HaloExchangeDev H = make_dev(this, d, 0);
  H.peerLL = d->d_peerLL[parity].p;
  for (int p = 0; p < kInlinePeers; ++p)
    H.peerWindowInline[p] = p < (int)d->h_peerLL[parity].size() ? d->h_peerLL[parity][p] : nullptr;
  H.myLL = (char*)d->arena + d->llOffset + (size_t)parity * d->llBytes;
  H.epoch32 = (unsigned)(d->epochLL & 0xffffffffull);
  GsRowsDev R = ogs->rows;
  R.nMasked = maskIds ? Nmasked : 0;
  R.maskIds = maskIds;
  return gs_exchange_ll_launch<T>(R, H, v, stream);
}
template int oogs_t::exchange_ll<double>(double*, int, gs_op, dlong, const dlong*, cudaStream_t);
template int oogs_t::exchange_ll<float>(float*, int, gs_op, dlong, const dlong*, cudaStream_t);

template int oogs_t::start<double>(double*, int, dlong, gs_op, cudaStream_t);
template int oogs_t::start<float>(float*, int, dlong, gs_op, cudaStream_t);
template int oogs_t::finish<double>(double*, int, dlong, gs_op, dlong, const dlong*, cudaStream_t);
template int oogs_t::finish<float>(float*, int, dlong, gs_op, dlong, const dlong*, cudaStream_t);

// ------------------------------------------------------------------------------------------ comm
// scalar all-reduce windows: slots [2][nranks][kMaxRed] doubles + flags [nranks], peer-mapped via IPC
int comm_setup_reduce(comm_t* c)
{
  if (c->nranks <= 1) return NRSB_OK;
  int rc;
  if (!c->h_err) {
    NRSB_CUDA(cudaHostAlloc((void**)&c->h_err, sizeof(int), cudaHostAllocMapped));
    *c->h_err = 0;
    NRSB_CUDA(cudaHostGetDevicePointer((void**)&c->d_err, c->h_err, 0));
  }
  const size_t slotDoubles = (size_t)2 * c->nranks * kMaxRed * 2;  // 16 bytes per value: flag-in-data words
  // one arena so that a single IPC handle covers slots and flags
  if ((rc = c->redSlots.alloc(slotDoubles + c->nranks))) return rc;
  if ((rc = c->redEpoch.alloc(1))) return rc;
  NRSB_CUDA(cudaDeviceSynchronize());
  std::vector<cudaIpcMemHandle_t> handles(c->nranks);
  NRSB_CUDA(cudaIpcGetMemHandle(&handles[c->rank], c->redSlots.p));
  c->allgather_bytes(handles.data(), sizeof(cudaIpcMemHandle_t));
  c->peerRedSlots.assign(c->nranks, nullptr);
  c->peerRedFlags.assign(c->nranks, nullptr);
  for (int p = 0; p < c->nranks; ++p) {
    void* base = c->redSlots.p;
    if (p != c->rank) NRSB_CUDA(cudaIpcOpenMemHandle(&base, handles[p], cudaIpcMemLazyEnablePeerAccess));
    c->peerRedSlots[p] = (double*)base;
    c->peerRedFlags[p] = (unsigned long long*)((double*)base + slotDoubles);
  }
  if ((rc = c->d_peerRedSlots.upload(c->peerRedSlots))) return rc;
  if ((rc = c->d_peerRedFlags.upload(c->peerRedFlags))) return rc;
  c->barrier();
  return NRSB_OK;
}

PeerReduce comm_t::peerReduce() const
{
  PeerReduce P;
  P.rank = rank;
  P.nranks = nranks;
  P.slots = d_peerRedSlots.p;
  P.flags = d_peerRedFlags.p;
  P.epoch = redEpoch.p;
  P.err = d_err;
  return P;
}

}  // namespace nrsb

// gs.hpp -- gather-scatter handles (ogs_t / oogs_t of 3rd_party/gslib/ogs/ogs.hpp:145-295).
#pragma once
#include <vector>

#include "common.cuh"

namespace nrsb {

// device view of the bucketed on-rank rows (see gs.cu header)
struct GsRowsDev {
  int nPairs = 0;
  const int2* pairs = nullptr;
  int nQuads = 0;
  const int4* quads = nullptr;
  int nOcts = 0;
  const int4* octs = nullptr;  // 2 x int4 per row
  int nGen = 0;
  const int* genStarts = nullptr;
  const int* genIds = nullptr;
  int nMasked = 0;  // optional Dirichlet ids zeroed in the same launch
  const int* maskIds = nullptr;
};

#ifdef __CUDACC__
// One row of the bucketed table, fetched BEFORE the data is ready (the table does not depend on it).
// n = number of copies (ids in id[0..n)), n = 1: a masked node (store zero), n = -1: general CSR row
// [id[0], id[1]) of genIds (more than 8 copies: not on a conforming hex mesh interior, kept for generality).
struct GsRowRef {
  int n;
  int id[8];
};

__device__ __forceinline__ GsRowRef gs_row_fetch(const GsRowsDev& R, long m)
{
  GsRowRef r;
  r.n = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) r.id[c] = 0;
  if (m < R.nPairs) {
    const int2 id = R.pairs[m];
    r.n = 2;
    r.id[0] = id.x;
    r.id[1] = id.y;
    return r;
  }
  m -= R.nPairs;
  if (m < R.nQuads) {
    const int4 id = R.quads[m];
    r.n = 4;
    r.id[0] = id.x;
    r.id[1] = id.y;
    r.id[2] = id.z;
    r.id[3] = id.w;
    return r;
  }
  m -= R.nQuads;
  if (m < R.nOcts) {
    const int4 ia = R.octs[2 * m], ib = R.octs[2 * m + 1];
    r.n = 8;
    r.id[0] = ia.x;
    r.id[1] = ia.y;
    r.id[2] = ia.z;
    r.id[3] = ia.w;
    r.id[4] = ib.x;
    r.id[5] = ib.y;
    r.id[6] = ib.z;
    r.id[7] = ib.w;
    return r;
  }
  m -= R.nOcts;
  if (m < R.nGen) {
    r.n = -1;
    r.id[0] = R.genStarts[m];
    r.id[1] = R.genStarts[m + 1];
    return r;
  }
  m -= R.nGen;
  if (m < R.nMasked) {
    r.n = 1;
    r.id[0] = R.maskIds[m];
  }
  return r;
}

#endif

// the on-rank gather-scatter + mask as phase 2 of the persistent axhelm launch (axhelm_tma.cu): every axhelm CTA
// arrives at a device-wide counter once its elements are stored, waits for the others, then takes its share of
// the bucketed rows.  Replaces the kernel boundary + second launch of ellipticOperator (ellipticOperator.cpp:158-168).
struct FusedRows {
  GsRowsDev R;
  unsigned long long* arrive = nullptr;  // monotone arrival counter of the axhelm CTAs
  unsigned long long target = 0;         // its value once every axhelm CTA of this launch has arrived
  // (any ax_tma launch) when set: CTA b stores  sum over its elements of  q_e^T A_e q_e  (energy form
  // lam0 * grad q . G grad q [+ lam1 * GwJ q^2], evaluated where the kernel already holds those factors) to
  // dotPartials[b].  For a continuous q (PCG's p) the sum over CTAs is  q^T Q^T A_L Q q = the weighted inner
  // product  sum invDegree * q * (Q Q^T A_L q)  that PCG.cpp:150-157 computes with a separate pass after the
  // gather-scatter.
  double* dotPartials = nullptr;
  // (any ax_tma launch) streamed gather-scatter: see AxDot::chunkDone (kernels.hpp)
  unsigned long long* chunkDone = nullptr;
  int chunkLen = 1;
};

template <typename T>
int gs_rows_launch(const GsRowsDev& R, int Nfields, dlong stride, T* q, cudaStream_t stream);
template <typename T>
int gs_csr_launch(dlong Ngather, int Nentries, dlong stride, const dlong* starts, const dlong* ids, T* q,
                  cudaStream_t stream);
template <typename T>
int mask_launch(dlong Nmasked, const dlong* maskIds, T* q, cudaStream_t stream);

// halo side (halo.cu)
struct HaloDev {
  int nRows = 0;                 // halo gather rows on this rank
  const int* rowStarts = nullptr;  // CSR over local copies of each halo row
  const int* rowIds = nullptr;
  const int* sendStarts = nullptr;  // CSR over (peer slot) destinations of each row
  const int* sendPeer = nullptr;    // peer index (0..nPeers-1)
  const int* sendSlot = nullptr;    // slot in that peer's receive window
  const int* recvStarts = nullptr;  // CSR over contributions (ascending global rank incl. self)
  const int* recvSrc = nullptr;     // -1 = own partial sum, else offset into the local receive buffer
};

struct SharedTopology {  // what the bootstrap layer (torch.distributed / MPI) discovered
  int rank = 0, nranks = 1;
  long nShared = 0;
  const hlong* sharedIds = nullptr;      // ascending
  const int* sharerOffsets = nullptr;    // nShared+1
  const int* sharerRanks = nullptr;      // ascending ranks per id, including this rank
};

class comm_t;

// ogs_t: setup result.  Host CSR arrays reproduce ogsSetup.cpp bit for bit (tests compare them).
class ogs_t {
 public:
  dlong N = 0;
  // --- reference-layout maps (host)
  dlong Nlocal = 0, NlocalGather = 0;  // nodes / rows that never leave the rank
  std::vector<dlong> localGatherOffsets, localGatherIds;
  dlong Nhalo = 0, NhaloGather = 0;
  std::vector<dlong> haloGatherOffsets, haloGatherIds;
  std::vector<hlong> haloBaseIds;  // global id of each halo row
  std::vector<double> invDegree;   // host copy
  // --- device
  double* d_invDegree = nullptr;
  float* d_invDegreePfloat = nullptr;
  dlong* d_localGatherOffsets = nullptr;  // full CSR (incl. singletons) for the kernel-level ABI
  dlong* d_localGatherIds = nullptr;
  int2* d_pairs = nullptr;
  int4* d_quads = nullptr;
  int4* d_octs = nullptr;
  int* d_genStarts = nullptr;
  int* d_genIds = nullptr;
  GsRowsDev rows;  // on-rank rows
  // halo rows: local part
  int* d_haloStarts = nullptr;
  int* d_haloIds = nullptr;

  int rank = 0, nranks = 1;
  // per halo row: sharer ranks (ascending, incl. self)
  std::vector<int> haloSharerOffsets, haloSharerRanks;

  ~ogs_t();
  int setup(dlong N, const hlong* ids, const SharedTopology* topo);
  int upload_inv_degree();
};

}  // namespace nrsb

// ------------------------------------------------------------------------------------------------
// Streamed gather-scatter (gs_stream.cu): the on-rank rows + Dirichlet mask of ellipticOperator
// (ellipticOperator.cpp:158-168) executed WHILE the persistent axhelm launch is still running, by a small
// co-resident kernel (one 128-thread block per SM in the registers the axhelm CTA leaves free, launched as a
// programmatic dependent launch).  The element list is cut into chunks of consecutive list positions; axhelm
// counts finished elements per chunk; a row becomes ready when the chunks of all its copies are complete.
// Rows are bucketed as in gs.cu, sorted by ready chunk (stable, so ascending base id inside a chunk) and cut into
// warp-sized units.  Sums are formed in the reference's order => bit-identical to gs_rows_kernel.
namespace nrsb {

struct GsStreamDev {
  int nUnits = 0;
  const int4* units = nullptr;  // {kind (0 pairs, 1 quads, 2 octs, 3 general, 4 mask), ready chunk, first row, rows}
  const int2* pairs = nullptr;
  const int4* quads = nullptr;
  const int4* octs = nullptr;
  const int* genStarts = nullptr;
  const int* genIds = nullptr;
  const int* maskIds = nullptr;
  int nChunks = 0, chunkLen = 1, Nelements = 0;
  const unsigned long long* done = nullptr;  // monotone per-chunk counters (never reset)
  unsigned long long epoch = 0;              // launches so far incl. this one: chunk c is complete at epoch * size(c)
  int withMask = 1;
  int* err = nullptr;  // host-mapped word: set to 1 when a wait timed out
};

class gs_stream_t {
 public:
  static constexpr int kMaxChunks = 32;
  static constexpr int kPairsPerLane = 6, kQuadsPerLane = 3, kMaskPerLane = 12;
  int Nelements = 0, chunkLen = 0, nChunks = 0;
  unsigned long long epoch = 0;
  int nUnits = 0;
  int2* d_pairs = nullptr;
  int4* d_quads = nullptr;
  int4* d_octs = nullptr;
  int* d_genStarts = nullptr;
  int* d_genIds = nullptr;
  int* d_maskIds = nullptr;
  int4* d_units = nullptr;
  unsigned long long* d_done = nullptr;
  int* h_err = nullptr;  // pinned + mapped
  int* d_err = nullptr;
  ~gs_stream_t();
  // rows: the (masked) handle's on-rank CSR; elementPos[e] = position of element e in the element list the axhelm
  // launch walks; nAx = number of axhelm CTAs of that launch
  int build(const ogs_t* ogs, const std::vector<dlong>& maskIds, const std::vector<dlong>& elementPos, int Np, int nAx);
  GsStreamDev dev(bool withMask) const;
  bool failed() const { return h_err && *h_err != 0; }
};

// launch behind the axhelm kernel on `stream` (programmatic dependent launch); S.epoch must already count it
template <typename T>
int gs_stream_launch(const GsStreamDev& S, T* q, cudaStream_t stream);
// can the stream kernel be co-resident with a CTA of `axRegs` registers x `axThreads` threads + `axSmem` bytes?
bool gs_stream_fits(int axRegsPerThread, int axThreads, size_t axSmemBytes);

}  // namespace nrsb

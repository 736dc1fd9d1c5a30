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
  int* err = nullptr;  // host-mapped error word: the device-wide wait of the launch gave up
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

// ogs.cpp -- host-side gather-scatter setup.
//
// Restates ogsSetup (3rd_party/gslib/ogs/src/ogsSetup.cpp:111-397) without gslib/MPI: the
// classification "which ids are shared with other ranks" (gslib gs_setup + crystal router in the
// reference, ogsSetup.cpp:150-175) is supplied by the bootstrap layer as a SharedTopology.
//
//  * ids == 0 are ignored (ogs.hpp:42-44).
//  * local rows: one per distinct id present only on this rank, ordered by the smallest local
//    index of the row; inside a row local indices ascend (ogsSetup.cpp:196-249).
//  * halo rows: ids shared with another rank, same ordering rule (the reference additionally
//    moves "owned" rows first, :272-349; ownership is irrelevant for the symmetric
//    gather-scatter, so it is not reproduced).
//  * invDegree[n] = 1 / (global multiplicity of node n), 1 for ignored nodes (:366-392).
#include <algorithm>
#include <cstdlib>
#include <numeric>

#include "gs.hpp"

namespace nrsb {

namespace {
struct node_t {
  hlong id;
  dlong idx;
};

template <typename T>
int upload(T** d, const std::vector<T>& h)
{
  *d = nullptr;
  if (h.empty()) return NRSB_OK;
  NRSB_CUDA(cudaMalloc((void**)d, h.size() * sizeof(T)));
  NRSB_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return NRSB_OK;
}

// rows (grouped by id) of `nodes` sorted by (id, idx); ordered by first idx
void build_rows(std::vector<node_t>& nodes, std::vector<dlong>& offsets, std::vector<dlong>& ids,
                std::vector<hlong>* baseIds)
{
  std::sort(nodes.begin(), nodes.end(),
            [](const node_t& a, const node_t& b) { return a.id != b.id ? a.id < b.id : a.idx < b.idx; });
  struct row_t {
    dlong first, start, count;
    hlong id;
  };
  std::vector<row_t> rows;
  for (size_t n = 0; n < nodes.size();) {
    size_t m = n;
    while (m < nodes.size() && nodes[m].id == nodes[n].id) ++m;
    rows.push_back({nodes[n].idx, (dlong)n, (dlong)(m - n), nodes[n].id});
    n = m;
  }
  std::sort(rows.begin(), rows.end(), [](const row_t& a, const row_t& b) { return a.first < b.first; });
  offsets.assign(rows.size() + 1, 0);
  ids.resize(nodes.size());
  if (baseIds) baseIds->resize(rows.size());
  dlong pos = 0;
  for (size_t r = 0; r < rows.size(); ++r) {
    offsets[r] = pos;
    for (dlong c = 0; c < rows[r].count; ++c) ids[pos++] = nodes[rows[r].start + c].idx;
    if (baseIds) (*baseIds)[r] = rows[r].id;
  }
  offsets[rows.size()] = pos;
}
}  // namespace

ogs_t::~ogs_t()
{
  cudaFree(d_invDegree);
  cudaFree(d_invDegreePfloat);
  cudaFree(d_localGatherOffsets);
  cudaFree(d_localGatherIds);
  cudaFree(d_pairs);
  cudaFree(d_quads);
  cudaFree(d_octs);
  cudaFree(d_genStarts);
  cudaFree(d_genIds);
  cudaFree(d_haloStarts);
  cudaFree(d_haloIds);
}

int ogs_t::setup(dlong N_, const hlong* ids, const SharedTopology* topo)
{
  N = N_;
  rank = topo ? topo->rank : 0;
  nranks = topo ? topo->nranks : 1;

  auto isShared = [&](hlong id) -> long {
    if (!topo || topo->nShared == 0) return -1;
    const hlong* b = topo->sharedIds;
    const hlong* e = b + topo->nShared;
    const hlong* p = std::lower_bound(b, e, id);
    return (p != e && *p == id) ? (long)(p - b) : -1;
  };

  std::vector<node_t> local, halo;
  local.reserve(N);
  for (dlong n = 0; n < N; ++n) {
    if (ids[n] == 0) continue;
    if (isShared(ids[n]) >= 0)
      halo.push_back({ids[n], n});
    else
      local.push_back({ids[n], n});
  }
  Nlocal = (dlong)local.size();
  Nhalo = (dlong)halo.size();
  build_rows(local, localGatherOffsets, localGatherIds, nullptr);
  NlocalGather = (dlong)localGatherOffsets.size() - 1;
  build_rows(halo, haloGatherOffsets, haloGatherIds, &haloBaseIds);
  NhaloGather = (dlong)haloGatherOffsets.size() - 1;

  // sharers and global multiplicity of halo rows
  haloSharerOffsets.assign(NhaloGather + 1, 0);
  haloSharerRanks.clear();
  for (dlong r = 0; r < NhaloGather; ++r) {
    const long s = isShared(haloBaseIds[r]);
    for (int c = topo->sharerOffsets[s]; c < topo->sharerOffsets[s + 1]; ++c)
      haloSharerRanks.push_back(topo->sharerRanks[c]);
    haloSharerOffsets[r + 1] = (int)haloSharerRanks.size();
  }

  // invDegree: local rows know their full multiplicity; halo rows need the remote counts,
  // which the caller fills in through set_halo_degree() after one integer exchange.  Until
  // then the local count is used (exact for single-rank runs).
  invDegree.assign(N, 1.0);
  for (dlong r = 0; r < NlocalGather; ++r) {
    const dlong cnt = localGatherOffsets[r + 1] - localGatherOffsets[r];
    for (dlong c = localGatherOffsets[r]; c < localGatherOffsets[r + 1]; ++c) invDegree[localGatherIds[c]] = 1.0 / cnt;
  }
  for (dlong r = 0; r < NhaloGather; ++r) {
    const dlong cnt = haloGatherOffsets[r + 1] - haloGatherOffsets[r];
    for (dlong c = haloGatherOffsets[r]; c < haloGatherOffsets[r + 1]; ++c) invDegree[haloGatherIds[c]] = 1.0 / cnt;
  }

  // ---- device layout: bucket on-rank rows by length
  std::vector<int2> pairs;
  std::vector<int4> quads, octs;
  std::vector<int> genStarts(1, 0), genIds;
  for (dlong r = 0; r < NlocalGather; ++r) {
    const dlong s = localGatherOffsets[r], cnt = localGatherOffsets[r + 1] - s;
    const dlong* g = &localGatherIds[s];
    if (cnt == 1) continue;
    if (cnt == 2)
      pairs.push_back(make_int2(g[0], g[1]));
    else if (cnt == 4)
      quads.push_back(make_int4(g[0], g[1], g[2], g[3]));
    else if (cnt == 8) {
      octs.push_back(make_int4(g[0], g[1], g[2], g[3]));
      octs.push_back(make_int4(g[4], g[5], g[6], g[7]));
    } else {
      for (dlong c = 0; c < cnt; ++c) genIds.push_back(g[c]);
      genStarts.push_back((int)genIds.size());
    }
  }
  // Row order inside a bucket: the reference's (ascending base id).  Measured (tools/gs_timing.py, E=4096):
  // sorting the buckets by the local index of the first or last copy changes the kernel time by < 2 %.
  int rc;
  if ((rc = upload(&d_pairs, pairs))) return rc;
  if ((rc = upload(&d_quads, quads))) return rc;
  if ((rc = upload(&d_octs, octs))) return rc;
  if (genStarts.size() > 1) {
    if ((rc = upload(&d_genStarts, genStarts))) return rc;
    if ((rc = upload(&d_genIds, genIds))) return rc;
  }
  rows.nPairs = (int)pairs.size();
  rows.pairs = d_pairs;
  rows.nQuads = (int)quads.size();
  rows.quads = d_quads;
  rows.nOcts = (int)octs.size() / 2;
  rows.octs = d_octs;
  rows.nGen = (int)genStarts.size() - 1;
  rows.genStarts = d_genStarts;
  rows.genIds = d_genIds;

  if ((rc = upload(&d_localGatherOffsets, localGatherOffsets))) return rc;
  if ((rc = upload(&d_localGatherIds, localGatherIds))) return rc;
  if (NhaloGather) {
    if ((rc = upload(&d_haloStarts, haloGatherOffsets))) return rc;
    if ((rc = upload(&d_haloIds, haloGatherIds))) return rc;
  }
  return upload_inv_degree();
}

int ogs_t::upload_inv_degree()
{
  cudaFree(d_invDegree);
  cudaFree(d_invDegreePfloat);
  d_invDegree = nullptr;
  d_invDegreePfloat = nullptr;
  int rc;
  if ((rc = upload(&d_invDegree, invDegree))) return rc;
  std::vector<float> f(invDegree.begin(), invDegree.end());
  return upload(&d_invDegreePfloat, f);
}

}  // namespace nrsb

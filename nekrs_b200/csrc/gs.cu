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

namespace nrsb {

// kRPT rows per thread.  All value loads of all rows of a thread are issued (predicated, no branches) before
// the first add, so a thread has up to 8 kRPT independent loads in flight.  Copies are summed in ascending
// local index, as the reference's CSR loop does (bit-identical sums).
// The kernel is launched as a programmatic dependent launch: blocks may become resident while the producer
// (axhelm) is still running, fetch their index entries, and sleep in pdl_wait() until its stores are visible.
template <typename T, int kRPT, int kBS>
__global__ void __launch_bounds__(kBS, kRPT == 1 ? 2048 / kBS : 1)  // 1 row per thread: 32 registers, 2048 threads per SM
    gs_rows_kernel(const GsRowsDev R, const int Nfields, const dlong stride, T* __restrict__ q)
{
  (void)Nfields;
  pdl_trigger();  // a persistent axhelm launch behind this kernel may start its prologue on drained SMs
  T* __restrict__ qf = q + (size_t)blockIdx.y * stride;
  GsRowRef r[kRPT];
#pragma unroll
  for (int j = 0; j < kRPT; ++j) r[j] = gs_row_fetch(R, ((long)blockIdx.x * kRPT + j) * blockDim.x + threadIdx.x);
  pdl_wait();
  T v[kRPT][8];
#pragma unroll
  for (int j = 0; j < kRPT; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) v[j][c] = (c < r[j].n && r[j].n > 1) ? qf[r[j].id[c]] : T(0);
#pragma unroll
  for (int j = 0; j < kRPT; ++j) {
    T s = T(0);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < r[j].n) s += v[j][c];
    if (r[j].n == 1) s = T(0);  // masked node
    if (r[j].n == -1) {
      for (int c = r[j].id[0]; c < r[j].id[1]; ++c) s += qf[R.genIds[c]];
      for (int c = r[j].id[0]; c < r[j].id[1]; ++c) qf[R.genIds[c]] = s;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < r[j].n) qf[r[j].id[c]] = s;
  }
}

int gs_rows_per_thread()
{
  static const int rpt = [] {
    const char* e = getenv("NRSB_GS_RPT");
    const int v = e ? atoi(e) : 1;
    return (v == 2 || v == 4) ? v : 1;
  }();
  return rpt;
}

template <typename T>
int gs_rows_launch(const GsRowsDev& R, int Nfields, dlong stride, T* q, cudaStream_t stream)
{
  const long total = (long)R.nPairs + R.nQuads + R.nOcts + R.nGen + R.nMasked;
  if (total == 0 || Nfields == 0) return NRSB_OK;
  const int rpt = gs_rows_per_thread();
  static const int bs = [] {
    const char* e = getenv("NRSB_GS_BS");
    const int v = e ? atoi(e) : 128;  // measured (tools/gs_timing.py): 64: 12.35, 128: 11.7-12.3, 256: 12.35, 512: 13.7, 1024: 17.4 us
    return (v == 64 || v == 256 || v == 512 || v == 1024) ? v : 128;
  }();
  const long perBlock = (long)bs * rpt;
  dim3 grid((unsigned)((total + perBlock - 1) / perBlock), Nfields);
  auto kern = rpt == 1 ? (bs == 64 ? gs_rows_kernel<T, 1, 64> : bs == 128 ? gs_rows_kernel<T, 1, 128>
                                    : (bs == 512 ? gs_rows_kernel<T, 1, 512>
                                                 : (bs == 1024 ? gs_rows_kernel<T, 1, 1024> : gs_rows_kernel<T, 1, 256>)))
                       : (rpt == 2 ? gs_rows_kernel<T, 2, 256> : gs_rows_kernel<T, 4, 256>);
  NRSB_CUDA(launch_pdl_consumer(kern, grid, dim3(rpt == 1 ? bs : 256), 0, stream, R, Nfields, stride, q));
  return NRSB_OK;
}
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

// axhelm_tma_nq.cu -- the persistent TMA-ring axhelm (axhelm_tma.inc) for even Nq other than 8.
// The bulk copies need 16-byte aligned slabs: Np * sizeof(T) is a multiple of 16 for even Nq only, so the odd Nq
// stay on the pencil kernels.  A consumer group is Nq^2 threads rounded up to whole warps (the padding lanes only
// take part in the barriers).  Ring shapes (consumer groups x stages) are what fits 227 KB of shared memory and the
// register file of one SM sub-partition (16 K registers per warp slot modulo 4):
//   fp64: Nq = 6: 6 x 12 (118-122 registers),  Nq = 10: 1 x 3 (255 registers; 2 x 2 = 9 warps caps a thread at 168
//         registers and spills 570 B: 52 us instead of 44 us)
//   fp32: Nq = 10: 4 x 4,  Nq = 12: 2 x 2
// Measured against the pencil kernels (~2 M nodes, B200, profiles/r2_axhelm_tma_nq.md): N=5 fp64 35.8 / 37.9 us,
// N=9 fp64 44.0 / 46.1 (Helmholtz 48.1 / 56.1), N=9 fp32 23.5 / 25.6, N=11 fp32 23.6 / 27.6 (Helmholtz 25.6 / 33.8).
// Not built because slower: Nq = 4 (50 us against 32: 512-byte q slabs, per-element barrier cost), Nq = 6 fp32 (tie).
#include "axhelm_tma.inc"

namespace nrsb {

bool ax_tma_nq_supported(int Nq, int precision)
{
  if (Nq == 10) return true;
  return (Nq == 6 && precision == 8) || (Nq == 12 && precision == 4);
}

template <typename T>
int ax_tma_nq_launch(int Nq, dlong Nelements, const dlong* elementList, const T* ggeo, const T* D_host,
                     const T* lambda0, const T* lambda1, int poisson, const T* q, T* Aq, cudaStream_t stream)
{
  if (Nelements == 0) return NRSB_OK;
#define NRSB_TMA(NQ, G, S_)                                                                                       \
  return poisson ? launch_tma<T, NQ, G, S_, true>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, q, Aq, \
                                                  stream)                                                        \
                 : launch_tma<T, NQ, G, S_, false>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, q, Aq, \
                                                   stream);
  if constexpr (sizeof(T) == 8) {
    if (Nq == 6) { NRSB_TMA(6, 6, 12) }
    if (Nq == 10) { NRSB_TMA(10, 1, 3) }
  } else {
    if (Nq == 10) { NRSB_TMA(10, 4, 4) }
    if (Nq == 12) { NRSB_TMA(12, 2, 2) }
  }
#undef NRSB_TMA
  set_last_error("axhelm TMA ring: Nq = " + std::to_string(Nq) + " is not built for this precision");
  return NRSB_ERR_INVALID;
}

template int ax_tma_nq_launch<double>(int, dlong, const dlong*, const double*, const double*, const double*,
                                      const double*, int, const double*, double*, cudaStream_t);
template int ax_tma_nq_launch<float>(int, dlong, const dlong*, const float*, const float*, const float*, const float*,
                                     int, const float*, float*, cudaStream_t);

}  // namespace nrsb

// axhelm_tma.cu -- entry points of the persistent TMA-ring axhelm for Nq = 8 (kernel: axhelm_tma.inc)
#include "axhelm_tma.inc"

namespace nrsb {

// variants 4..6 of the axhelm dispatch (Nq = 8, constant coefficients only)
template <typename T>
int ax_tma_launch(int Nq, int variant, dlong Nelements, const dlong* elementList, const T* ggeo, const T* D_host,
                  const T* lambda0, const T* lambda1, int poisson, const T* q, T* Aq, cudaStream_t stream, AxDot* dot)
{
  if (Nelements == 0) return NRSB_OK;
  if (Nq != 8) {
    set_last_error("axhelm variants 4-6 (TMA ring) are built for Nq = 8 only");
    return NRSB_ERR_INVALID;
  }
#define NRSB_TMA(G, S_)                                                                                              \
  return poisson ? launch_tma<T, 8, G, S_, true>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, q, Aq, stream, \
                                                 nullptr, nullptr, dot)                                                 \
                 : launch_tma<T, 8, G, S_, false>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, q, Aq, stream, \
                                                  nullptr, nullptr, dot);
  // NSTAGES must be a multiple of NGROUPS: group g then only ever touches stages == g (mod NGROUPS),
  // i.e. it owns a private sub-ring, and every mbarrier wait is at most one phase behind.
  // Helmholtz stages carry a seventh plane (GwJ): 6 x 32 KB + work buffers exceed the 227 KB of an SM for the
  // 3x6 / 5x5 (fp64) and 6x12 / 8x8 (fp32) rings, so the Helmholtz operator always takes the 4-stage ring.
  if (!poisson) variant = 4;
  if (sizeof(T) == 8) {
    if (variant == 4) { NRSB_TMA(4, 4) }
    if (variant == 5) {
      if (dot && dot->partials)
        return poisson ? launch_tma<T, 8, 3, 6, true, false, false, true>(Nelements, elementList, ggeo, D_host, lambda0,
                                                                          lambda1, q, Aq, stream, nullptr, nullptr, dot)
                       : launch_tma<T, 8, 3, 6, false, false, false, true>(Nelements, elementList, ggeo, D_host,
                                                                           lambda0, lambda1, q, Aq, stream, nullptr,
                                                                           nullptr, dot);
      NRSB_TMA(3, 6)
    }
    NRSB_TMA(5, 5)
  } else {
    if (variant == 4) { NRSB_TMA(4, 8) }
    if (variant == 5) { NRSB_TMA(6, 12) }
    NRSB_TMA(8, 8)
  }
#undef NRSB_TMA
}

// Ax over [halo elements..., interior elements...] with the halo push done inside the launch (kFused)
template <typename T>
int ax_tma_fused_launch(int Nq, int variant, dlong Nelements, const dlong* elementList, const T* ggeo, const T* D_host,
                        const T* lambda0, const T* lambda1, int poisson, const T* q, T* Aq, const FusedHalo& F,
                        cudaStream_t stream, AxDot* dot)
{
  if (Nelements == 0) return NRSB_OK;
  if (Nq != 8) {
    set_last_error("fused axhelm + halo push is built for Nq = 8 only");
    return NRSB_ERR_INVALID;
  }
  (void)variant;
#define NRSB_TMAF(G, S_)                                                                                      \
  return poisson ? launch_tma<T, 8, G, S_, true, true>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, q, \
                                                       Aq, stream, &F, nullptr, dot)                           \
                 : launch_tma<T, 8, G, S_, false, true>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, q, \
                                                        Aq, stream, &F, nullptr, dot);
  if (sizeof(T) == 8) {
    if (dot && dot->partials)
      return poisson ? launch_tma<T, 8, 3, 6, true, true, false, true>(Nelements, elementList, ggeo, D_host, lambda0,
                                                                       lambda1, q, Aq, stream, &F, nullptr, dot)
                     : launch_tma<T, 8, 3, 6, false, true, false, true>(Nelements, elementList, ggeo, D_host, lambda0,
                                                                        lambda1, q, Aq, stream, &F, nullptr, dot);
    NRSB_TMAF(3, 6)
  }
  NRSB_TMAF(6, 12)
#undef NRSB_TMAF
}
// Ax + mask + on-rank gather-scatter in ONE launch (single rank: the whole ellipticOperator; several ranks:
// halo rows are pushed by the pusher CTAs and folded by oogs::finish).  rows->target is advanced.
template <typename T>
int ax_tma_gs_launch(int Nq, dlong Nelements, const dlong* elementList, const T* ggeo, const T* D_host,
                     const T* lambda0, const T* lambda1, int poisson, const T* q, T* Aq, const FusedHalo* F,
                     FusedRows* rows, cudaStream_t stream, AxDot* dot)
{
  if (Nelements == 0) return NRSB_OK;
  if (Nq != 8) {
    set_last_error("fused axhelm + gather-scatter is built for Nq = 8 only");
    return NRSB_ERR_INVALID;
  }
#define NRSB_TMAG(G_, S_, FUSED)                                                                                   \
  return poisson ? launch_tma<T, 8, G_, S_, true, FUSED, true>(Nelements, elementList, ggeo, D_host, lambda0, lambda1, \
                                                              q, Aq, stream, F, rows, dot)                           \
                 : launch_tma<T, 8, G_, S_, false, FUSED, true>(Nelements, elementList, ggeo, D_host, lambda0,       \
                                                               lambda1, q, Aq, stream, F, rows, dot);
  if (F) {
    if (sizeof(T) == 8) { NRSB_TMAG(3, 6, true) }
    NRSB_TMAG(6, 12, true)
  }
  if (sizeof(T) == 8) { NRSB_TMAG(3, 6, false) }
  NRSB_TMAG(6, 12, false)
#undef NRSB_TMAG
}
template int ax_tma_gs_launch<double>(int, dlong, const dlong*, const double*, const double*, const double*,
                                      const double*, int, const double*, double*, const FusedHalo*, FusedRows*,
                                      cudaStream_t, AxDot*);
template int ax_tma_gs_launch<float>(int, dlong, const dlong*, const float*, const float*, const float*, const float*,
                                     int, const float*, float*, const FusedHalo*, FusedRows*, cudaStream_t, AxDot*);

template int ax_tma_fused_launch<double>(int, int, dlong, const dlong*, const double*, const double*, const double*,
                                         const double*, int, const double*, double*, const FusedHalo&, cudaStream_t,
                                         AxDot*);
template int ax_tma_fused_launch<float>(int, int, dlong, const dlong*, const float*, const float*, const float*,
                                        const float*, int, const float*, float*, const FusedHalo&, cudaStream_t,
                                        AxDot*);

template int ax_tma_launch<double>(int, int, dlong, const dlong*, const double*, const double*, const double*,
                                   const double*, int, const double*, double*, cudaStream_t, AxDot*);
template int ax_tma_launch<float>(int, int, dlong, const dlong*, const float*, const float*, const float*,
                                  const float*, int, const float*, float*, cudaStream_t, AxDot*);

}  // namespace nrsb

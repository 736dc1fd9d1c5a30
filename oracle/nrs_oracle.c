/*
 * nrs_oracle.c -- CPU restatement of the nekRS v23.0 pressure-Poisson kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (nekrs_b200/, include/) may
 * link, import or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker /
 * CPU baseline.
 *
 * Each function restates ONE reference kernel with the SAME floating point
 * operation order as the reference's SERIAL (.c) backend (or, where the
 * reference only ships OKL, the order of one OKL thread), with the polynomial
 * order as a run-time argument instead of the reference's compile-time p_Nq.
 * Parity pin: tests/test_oracle_vs_ref.py checks every function that has a
 * reference `.c` bit-for-bit against that file compiled from /root/reference
 * (oracle/build_ref.py -> oracle/_ref/).
 *
 * Build: gcc -O2 -fPIC -shared -std=c99 (no -ffast-math, no -march: keeps
 * the arithmetic identical to the reference CI build, .github/workflows/ci.yml:19).
 *
 * Type suffixes: _d = dfloat (double), _f = pfloat (float).  dlong = int.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef int dlong;

/* ------------------------------------------------------------------------- */
/* ellipticPartialAxCoeffHex3D_v0   kernels/elliptic/ellipticPartialAxCoeffHex3D.c:1-98
 * ggeo planes: G00,G01,G11,G12,G02,G22,GWJ = 0..6 (src/mesh/mesh3D.h:96-102)      */
#define DEF_AX(T, SUF)                                                                         \
  void orc_ax_##SUF(dlong Nelements, dlong offset, dlong loffset, const dlong *elementList,    \
                    const T *ggeo, const T *D, const T *S, const T *lambda0, const T *lambda1, \
                    const T *q, T *Aq, int Nq, int poisson, int p_lambda)                      \
  {                                                                                            \
    const int Np = Nq * Nq * Nq;                                                               \
    T *s_q = (T *)malloc(4 * (size_t)Np * sizeof(T));                                          \
    T *s_Gqr = s_q + Np, *s_Gqs = s_q + 2 * Np, *s_Gqt = s_q + 3 * Np;                         \
    (void)offset;                                                                              \
    for (dlong e = 0; e < Nelements; ++e) {                                                    \
      const dlong element = elementList[e];                                                    \
      for (int n = 0; n < Np; ++n)                                                             \
        s_q[n] = q[n + element * Np];                                                          \
      for (int k = 0; k < Nq; ++k)                                                             \
        for (int j = 0; j < Nq; ++j)                                                           \
          for (int i = 0; i < Nq; ++i) {                                                       \
            const int n = k * Nq * Nq + j * Nq + i;                                            \
            const dlong gbase = element * 7 * Np + n;                                          \
            const T r_G00 = ggeo[gbase + 0 * Np];                                              \
            const T r_G01 = ggeo[gbase + 1 * Np];                                              \
            const T r_G11 = ggeo[gbase + 2 * Np];                                              \
            const T r_G12 = ggeo[gbase + 3 * Np];                                              \
            const T r_G02 = ggeo[gbase + 4 * Np];                                              \
            const T r_G22 = ggeo[gbase + 5 * Np];                                              \
            const dlong id = element * Np + n;                                                 \
            const T r_lam0 = lambda0[p_lambda * id + 0 * loffset];                             \
            T qr = 0, qs = 0, qt = 0;                                                          \
            for (int m = 0; m < Nq; m++) {                                                     \
              qr += S[m * Nq + i] * s_q[k * Nq * Nq + j * Nq + m];                             \
              qs += S[m * Nq + j] * s_q[k * Nq * Nq + m * Nq + i];                             \
              qt += S[m * Nq + k] * s_q[m * Nq * Nq + j * Nq + i];                             \
            }                                                                                  \
            T Gqr = r_G00 * qr;                                                                \
            Gqr += r_G01 * qs;                                                                 \
            Gqr += r_G02 * qt;                                                                 \
            T Gqs = r_G01 * qr;                                                                \
            Gqs += r_G11 * qs;                                                                 \
            Gqs += r_G12 * qt;                                                                 \
            T Gqt = r_G02 * qr;                                                                \
            Gqt += r_G12 * qs;                                                                 \
            Gqt += r_G22 * qt;                                                                 \
            s_Gqr[n] = r_lam0 * Gqr;                                                           \
            s_Gqs[n] = r_lam0 * Gqs;                                                           \
            s_Gqt[n] = r_lam0 * Gqt;                                                           \
          }                                                                                    \
      for (int k = 0; k < Nq; k++)                                                             \
        for (int j = 0; j < Nq; ++j)                                                           \
          for (int i = 0; i < Nq; ++i) {                                                       \
            const int n = k * Nq * Nq + j * Nq + i;                                            \
            const dlong gbase = element * 7 * Np + n;                                          \
            const dlong id = element * Np + n;                                                 \
            T r_Aq = 0;                                                                        \
            if (!poisson) {                                                                    \
              const T r_lam1 = lambda1[p_lambda * id + 0 * loffset];                           \
              r_Aq = ggeo[gbase + 6 * Np] * r_lam1 * s_q[n];                                   \
            }                                                                                  \
            T r_Aqr = 0, r_Aqs = 0, r_Aqt = 0;                                                 \
            for (int m = 0; m < Nq; m++) {                                                     \
              r_Aqr += D[m * Nq + i] * s_Gqr[k * Nq * Nq + j * Nq + m];                        \
              r_Aqs += D[m * Nq + j] * s_Gqs[k * Nq * Nq + m * Nq + i];                        \
              r_Aqt += D[m * Nq + k] * s_Gqt[m * Nq * Nq + j * Nq + i];                        \
            }                                                                                  \
            Aq[id] = r_Aqr + r_Aqs + r_Aqt + r_Aq;                                             \
          }                                                                                    \
    }                                                                                          \
    free(s_q);                                                                                 \
  }
DEF_AX(double, d)
DEF_AX(float, f)

/* ------------------------------------------------------------------------- */
/* ellipticStressPartialAxCoeffHex3D_v0   kernels/elliptic/ellipticStressPartialAxCoeffHex3D.c:1-169
 * three coupled fields (u, v, w `offset` apart), viscous stress form
 *     A q = -div( lambda0 (grad q + grad q^T) ) + lambda1 q       (weak form, collocated JW)
 * vgeo planes: rx,ry,rz,sx,sy,sz,tx,ty,tz = 0..8, J = 9, JW = 10, 1/JW = 11 (src/mesh/mesh3D.h:82-93),
 * coefficients per field `loffset` apart, per node when p_lambda = 1.
 * Restated as three passes over the element (reference derivatives -> stress fluxes -> weak divergence); every
 * scalar is accumulated in the reference's order (m ascending; the three directions of the divergence interleaved
 * in ONE accumulator, :139-151), so the result has the reference's bits at -O2.                                 */
#define DEF_AX_STRESS(T, SUF)                                                                        \
  void orc_ax_stress_##SUF(dlong Nelements, dlong offset, dlong loffset, const dlong *elementList,   \
                           const T *vgeo, const T *D, const T *lambda0, const T *lambda1, const T *q, \
                           T *Aq, int Nq, int p_lambda)                                              \
  {                                                                                                  \
    const int Np = Nq * Nq * Nq, Nq2 = Nq * Nq;                                                      \
    T *fld = (T *)malloc(12 * (size_t)Np * sizeof(T)); /* 3 fields + 9 fluxes */                     \
    T *flux = fld + 3 * Np;                            /* [field][direction r,s,t][Np] */            \
    for (dlong el = 0; el < Nelements; ++el) {                                                       \
      const dlong e = elementList[el];                                                               \
      const T *g = vgeo + (size_t)e * Np * 12;                                                       \
      for (int f = 0; f < 3; ++f)                                                                    \
        for (int n = 0; n < Np; ++n)                                                                 \
          fld[f * Np + n] = q[(size_t)e * Np + n + (size_t)f * offset];                              \
      for (int k = 0; k < Nq; ++k)                                                                   \
        for (int j = 0; j < Nq; ++j)                                                                 \
          for (int i = 0; i < Nq; ++i) {                                                             \
            const int n = k * Nq2 + j * Nq + i;                                                      \
            T dr[3], ds[3], dt[3]; /* reference derivatives of the three fields */                   \
            for (int f = 0; f < 3; ++f) {                                                            \
              const T *s = fld + f * Np;                                                             \
              T a = 0, b = 0, c = 0;                                                                 \
              for (int m = 0; m < Nq; ++m) {                                                         \
                a += D[i * Nq + m] * s[k * Nq2 + j * Nq + m];                                        \
                b += D[j * Nq + m] * s[k * Nq2 + m * Nq + i];                                        \
                c += D[k * Nq + m] * s[m * Nq2 + j * Nq + i];                                        \
              }                                                                                      \
              dr[f] = a;                                                                             \
              ds[f] = b;                                                                             \
              dt[f] = c;                                                                             \
            }                                                                                        \
            const T rx = g[0 * Np + n], ry = g[1 * Np + n], rz = g[2 * Np + n];                      \
            const T sx = g[3 * Np + n], sy = g[4 * Np + n], sz = g[5 * Np + n];                      \
            const T tx = g[6 * Np + n], ty = g[7 * Np + n], tz = g[8 * Np + n];                      \
            const T JW = g[10 * Np + n];                                                             \
            T grad[3][3]; /* grad[f][x,y,z] */                                                       \
            for (int f = 0; f < 3; ++f) {                                                            \
              grad[f][0] = rx * dr[f] + sx * ds[f] + tx * dt[f];                                     \
              grad[f][1] = ry * dr[f] + sy * ds[f] + ty * dt[f];                                     \
              grad[f][2] = rz * dr[f] + sz * ds[f] + tz * dt[f];                                     \
            }                                                                                        \
            const dlong id = e * Np + n;                                                             \
            for (int f = 0; f < 3; ++f) {                                                            \
              const T lam0 = lambda0[p_lambda * id + f * loffset];                                   \
              const T s1 = lam0 * JW * (grad[f][0] + grad[0][f]);                                    \
              const T s2 = lam0 * JW * (grad[f][1] + grad[1][f]);                                    \
              const T s3 = lam0 * JW * (grad[f][2] + grad[2][f]);                                    \
              flux[(3 * f + 0) * Np + n] = rx * s1 + ry * s2 + rz * s3;                              \
              flux[(3 * f + 1) * Np + n] = sx * s1 + sy * s2 + sz * s3;                              \
              flux[(3 * f + 2) * Np + n] = tx * s1 + ty * s2 + tz * s3;                              \
            }                                                                                        \
          }                                                                                          \
      for (int k = 0; k < Nq; ++k)                                                                   \
        for (int j = 0; j < Nq; ++j)                                                                 \
          for (int i = 0; i < Nq; ++i) {                                                             \
            const int n = k * Nq2 + j * Nq + i;                                                      \
            const dlong id = e * Np + n;                                                             \
            const T JW = g[10 * Np + n];                                                             \
            for (int f = 0; f < 3; ++f) {                                                            \
              const T *Fr = flux + (3 * f + 0) * Np, *Fs = flux + (3 * f + 1) * Np,                  \
                      *Ft = flux + (3 * f + 2) * Np;                                                 \
              T acc = 0;                                                                             \
              for (int m = 0; m < Nq; ++m) {                                                         \
                acc += D[m * Nq + i] * Fr[k * Nq2 + j * Nq + m];                                     \
                acc += D[m * Nq + j] * Fs[k * Nq2 + m * Nq + i];                                     \
                acc += D[m * Nq + k] * Ft[m * Nq2 + j * Nq + i];                                     \
              }                                                                                      \
              const T lam1 = lambda1[p_lambda * id + f * loffset];                                   \
              Aq[id + (size_t)f * offset] = acc + lam1 * JW * fld[f * Np + n];                       \
            }                                                                                        \
          }                                                                                          \
    }                                                                                                \
    free(fld);                                                                                       \
  }
DEF_AX_STRESS(double, d)
DEF_AX_STRESS(float, f)

/* ------------------------------------------------------------------------- */
/* mask   kernels/core/mask.okl : q[maskIds[n]] = 0                          */
#define DEF_MASK(T, SUF)                                            \
  void orc_mask_##SUF(dlong Nmasked, const dlong *maskIds, T *q)    \
  {                                                                 \
    for (dlong n = 0; n < Nmasked; ++n)                             \
      q[maskIds[n]] = 0;                                            \
  }
DEF_MASK(double, d)
DEF_MASK(float, f)

/* ------------------------------------------------------------------------- */
/* gatherScatterMany_<T>_add  3rd_party/gslib/ogs/okl/gatherScatterMany.okl:
 * one OKL thread per gather row: sum the row's copies in CSR order (start ..
 * end), write the sum back to every copy.  k fields with stride.            */
#define DEF_GS(T, SUF)                                                                       \
  void orc_gs_add_##SUF(dlong Ngather, const dlong *starts, const dlong *ids, int k,         \
                        dlong stride, T *q)                                                  \
  {                                                                                          \
    for (int fld = 0; fld < k; ++fld)                                                        \
      for (dlong g = 0; g < Ngather; ++g) {                                                  \
        const dlong start = starts[g], end = starts[g + 1];                                  \
        if (start + 1 == end)                                                                \
          continue; /* singleton rows are skipped, gatherScatterMany.okl */                  \
        T gq = 0;                                                                            \
        for (dlong n = start; n < end; ++n)                                                  \
          gq += q[ids[n] + fld * stride];                                                    \
        for (dlong n = start; n < end; ++n)                                                  \
          q[ids[n] + fld * stride] = gq;                                                     \
      }                                                                                      \
  }
DEF_GS(double, d)
DEF_GS(float, f)

void orc_gs_min_i(dlong Ngather, const dlong *starts, const dlong *ids, int *q)
{
  for (dlong g = 0; g < Ngather; ++g) {
    const dlong start = starts[g], end = starts[g + 1];
    int gq = q[ids[start]];
    for (dlong n = start + 1; n < end; ++n)
      gq = q[ids[n]] < gq ? q[ids[n]] : gq;
    for (dlong n = start; n < end; ++n)
      q[ids[n]] = gq;
  }
}

/* ------------------------------------------------------------------------- */
/* ellipticBlockUpdatePCG  kernels/elliptic/ellipticBlockUpdatePCG.c:27-53   */
void orc_update_pcg_d(dlong N, dlong offset, int Nfields, const double *invDegree, const double *Ap,
                      double alpha, double *r, double *rdotr_out)
{
  double rdotr = 0;
  for (int fld = 0; fld < Nfields; fld++)
    for (int i = 0; i < N; ++i) {
      const dlong n = i + fld * offset;
      const double rn = r[n] - alpha * Ap[n];
      rdotr += rn * rn * invDegree[i];
      r[n] = rn;
    }
  rdotr_out[0] = rdotr;
}

/* ------------------------------------------------------------------------- */
/* linAlg serial kernels  kernels/linAlg/{axpbyMany,axpby,axmyz,axmy,...}.c and
 * the OKL-only ones (fill, scaleMany, add, adyMany, sum).                   */
#define DEF_LINALG(T, SUF)                                                                         \
  /* y = a x + b y   axpbyMany.c */                                                                \
  void orc_axpby_many_##SUF(dlong N, int Nfields, dlong offset, T a, const T *x, T b, T *y)        \
  {                                                                                                \
    for (int fld = 0; fld < Nfields; fld++)                                                        \
      for (dlong n = 0; n < N; ++n) {                                                              \
        const dlong id = n + fld * offset;                                                         \
        y[id] = a * x[id] + b * y[id];                                                             \
      }                                                                                            \
  }                                                                                                \
  /* z = a x .* y  axmyz.c */                                                                      \
  void orc_axmyz_##SUF(dlong N, T a, const T *x, const T *y, T *z)                                 \
  {                                                                                                \
    for (dlong n = 0; n < N; ++n)                                                                  \
      z[n] = a * x[n] * y[n];                                                                      \
  }                                                                                                \
  /* y = a x .* y  axmy.c */                                                                       \
  void orc_axmy_##SUF(dlong N, T a, const T *x, T *y)                                              \
  {                                                                                                \
    for (dlong n = 0; n < N; ++n)                                                                  \
      y[n] = a * y[n] * x[n]; /* alpha*ai*wi, axmy.c */                                           \
  }                                                                                                \
  void orc_fill_##SUF(dlong N, T a, T *x)                                                          \
  {                                                                                                \
    for (dlong n = 0; n < N; ++n)                                                                  \
      x[n] = a;                                                                                    \
  }                                                                                                \
  void orc_add_##SUF(dlong N, T a, T *x)                                                           \
  {                                                                                                \
    for (dlong n = 0; n < N; ++n)                                                                  \
      x[n] += a;                                                                                   \
  }                                                                                                \
  void orc_scale_many_##SUF(dlong N, int Nfields, dlong offset, T a, T *x)                         \
  {                                                                                                \
    for (int fld = 0; fld < Nfields; fld++)                                                        \
      for (dlong n = 0; n < N; ++n)                                                                \
        x[n + fld * offset] *= a;                                                                  \
  }                                                                                                \
  /* weightedInnerProdMany.c : sum_fld sum_n w[n] x y */                                           \
  T orc_weighted_inner_prod_many_##SUF(dlong N, int Nfields, dlong offset, const T *w, const T *x, \
                                       const T *y)                                                 \
  {                                                                                                \
    T wxy = 0;                                                                                     \
    for (int fld = 0; fld < Nfields; fld++)                                                        \
      for (dlong n = 0; n < N; ++n) {                                                              \
        const dlong id = n + fld * offset;                                                         \
        wxy += x[id] * y[id] * w[n]; /* ai*bi*wi */                                                \
      }                                                                                            \
    return wxy;                                                                                    \
  }                                                                                                \
  /* weightedNorm2Many.c : sum w x^2 (sqrt taken by linAlg.cpp) */                                 \
  T orc_weighted_norm2_many_##SUF(dlong N, int Nfields, dlong offset, const T *w, const T *x)      \
  {                                                                                                \
    T wx2 = 0;                                                                                     \
    for (int fld = 0; fld < Nfields; fld++)                                                        \
      for (dlong n = 0; n < N; ++n) {                                                              \
        const dlong id = n + fld * offset;                                                         \
        wx2 += x[id] * x[id] * w[n];                                                               \
      }                                                                                            \
    return wx2;                                                                                    \
  }                                                                                                \
  T orc_sum_##SUF(dlong N, const T *x)                                                             \
  {                                                                                                \
    T s = 0;                                                                                       \
    for (dlong n = 0; n < N; ++n)                                                                  \
      s += x[n];                                                                                   \
    return s;                                                                                      \
  }                                                                                                \
  T orc_inner_prod_##SUF(dlong N, const T *x, const T *y)                                          \
  {                                                                                                \
    T s = 0;                                                                                       \
    for (dlong n = 0; n < N; ++n)                                                                  \
      s += x[n] * y[n];                                                                            \
    return s;                                                                                      \
  }
DEF_LINALG(double, d)
DEF_LINALG(float, f)

/* weightedInnerProdMulti.okl : out[k] = sum_n w[n] x[n + k*offset] y[n], k < NVec */
void orc_weighted_inner_prod_multi_d(dlong N, int NVec, dlong offset, const double *w, const double *x,
                                     const double *y, double *out)
{
  for (int k = 0; k < NVec; ++k) {
    double s = 0;
    for (dlong n = 0; n < N; ++n)
      s += w[n] * x[n + (size_t)k * offset] * y[n];
    out[k] = s;
  }
}

/* copyDfloatToPfloat.c / copyPfloatToDfloat.c */
void orc_copy_d2f(dlong N, const double *x, float *y)
{
  for (dlong n = 0; n < N; ++n)
    y[n] = x[n];
}
void orc_copy_f2d(dlong N, const float *x, double *y)
{
  for (dlong n = 0; n < N; ++n)
    y[n] = x[n];
}

/* axmyzManyPfloat.c:  z(double) = double(alpha(float) * x(double) * y(float)) */
void orc_axmyz_many_pfloat(dlong N, int Nfields, dlong offset, float alpha, const double *x, const float *y,
                           double *z)
{
  for (dlong n = 0; n < N; ++n)
    for (int fld = 0; fld < Nfields; ++fld) {
      const int id = n + fld * offset;
      z[id] = (double)(alpha * x[id] * y[id]);
    }
}

/* ------------------------------------------------------------------------- */
/* updateChebyshev.okl:  x += d ; r -= SAd ; d = dCoeff d + rCoeff r
 * updateFourthKindChebyshev.okl:  x += beta d ; r -= Ad                      */
void orc_update_chebyshev_f(dlong N, float dCoeff, float rCoeff, const float *SAd, float *d, float *r, float *x)
{
  for (dlong n = 0; n < N; ++n) {
    const float dn = d[n];
    const float rnp1 = r[n] - SAd[n];
    x[n] = x[n] + dn;
    r[n] = rnp1;
    d[n] = dCoeff * dn + rCoeff * rnp1;
  }
}
void orc_update_fourth_chebyshev_f(dlong N, float beta, const float *Ad, const float *d, float *r, float *x)
{
  for (dlong n = 0; n < N; ++n) {
    x[n] = x[n] + beta * d[n];
    r[n] = r[n] - Ad[n];
  }
}

/* ------------------------------------------------------------------------- */
/* preFDM  kernels/elliptic/preFDM.c:15-93                                   */
#define IDXE(k, j, i) ((k) * Nqe * Nqe + (j) * Nqe + (i))
void orc_pre_fdm_f(dlong Nelements, const float *u, float *work1, int Nq)
{
  const int Nqe = Nq + 2, Npe = Nqe * Nqe * Nqe, Np = Nq * Nq * Nq;
  for (dlong e = 0; e < Nelements; ++e) {
    float *w = work1 + (size_t)e * Npe;
    const float *ue = u + (size_t)e * Np;
#define U(k, j, i) ue[((k)-1) * Nq * Nq + ((j)-1) * Nq + ((i)-1)]
    for (int k = 0; k < Nqe; ++k)
      for (int j = 0; j < Nqe; ++j)
        for (int i = 0; i < Nqe; ++i) {
          const int in = i >= 1 && i < Nqe - 1 && j >= 1 && j < Nqe - 1 && k >= 1 && k < Nqe - 1;
          w[IDXE(k, j, i)] = in ? U(k, j, i) : 0.0f;
        }
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b) {
        w[IDXE(0, a, b)] = U(2, a, b);
        w[IDXE(Nqe - 1, a, b)] = U(Nqe - 3, a, b);
        w[IDXE(a, 0, b)] = U(a, 2, b);
        w[IDXE(a, Nqe - 1, b)] = U(a, Nqe - 3, b);
        w[IDXE(a, b, 0)] = U(a, b, 2);
        w[IDXE(a, b, Nqe - 1)] = U(a, b, Nqe - 3);
      }
#undef U
  }
}

/* fusedFDM  kernels/elliptic/fusedFDM.c:3-224.  restrict=1: RAS (Su is Nq^3 per
 * element, multiplied by wts); restrict=0: ASM (Su is (Nq+2)^3, u gets the
 * overlap planes).  NB the serial reference ignores elementList (fusedFDM.c:31:
 * `element = my_elem`); so does this restatement, the argument is kept for
 * signature parity.                                                          */
void orc_fused_fdm_f(dlong Nelements, const dlong *elementList, float *Su, const float *S_x, const float *S_y,
                     const float *S_z, const float *inv_L, const float *wts, float *u, int Nq, int p_restrict)
{
  const int Nqe = Nq + 2, Npe = Nqe * Nqe * Nqe;
  float *Sx = (float *)malloc(sizeof(float) * (3 * Nqe * Nqe + 2 * (size_t)Npe));
  float *Sy = Sx + Nqe * Nqe, *Sz = Sy + Nqe * Nqe;
  float *tmp = Sz + Nqe * Nqe, *work2 = tmp + Npe;
  (void)elementList;
  for (dlong elem = 0; elem < Nelements; ++elem) {
    float *w1 = u + (size_t)elem * Npe;
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        w1[IDXE(0, a, b)] = w1[IDXE(0, a, b)] - w1[IDXE(2, a, b)];
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        w1[IDXE(Nqe - 1, a, b)] = w1[IDXE(Nqe - 1, a, b)] - w1[IDXE(Nqe - 3, a, b)];
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        w1[IDXE(a, 0, b)] = w1[IDXE(a, 0, b)] - w1[IDXE(a, 2, b)];
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        w1[IDXE(a, Nqe - 1, b)] = w1[IDXE(a, Nqe - 1, b)] - w1[IDXE(a, Nqe - 3, b)];
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        w1[IDXE(a, b, 0)] = w1[IDXE(a, b, 0)] - w1[IDXE(a, b, 2)];
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        w1[IDXE(a, b, Nqe - 1)] = w1[IDXE(a, b, Nqe - 1)] - w1[IDXE(a, b, Nqe - 3)];

    for (int n = 0; n < Nqe * Nqe; ++n) {
      Sx[n] = S_x[n + (size_t)elem * Nqe * Nqe];
      Sy[n] = S_y[n + (size_t)elem * Nqe * Nqe];
      Sz[n] = S_z[n + (size_t)elem * Nqe * Nqe];
    }
#define SX(a, b) Sx[(a) * Nqe + (b)]
#define SY(a, b) Sy[(a) * Nqe + (b)]
#define SZ(a, b) Sz[(a) * Nqe + (b)]
    /* work2[k][i][j] = sum_l Sx[l][j] u[k][i][l] */
    for (int k = 0; k < Nqe; k++)
      for (int j = 0; j < Nqe; j++)
        for (int i = 0; i < Nqe; i++) {
          float value = 0.0f;
          for (int l = 0; l < Nqe; l++)
            value += SX(l, j) * w1[IDXE(k, i, l)];
          work2[IDXE(k, i, j)] = value;
        }
    /* tmp[j][k][i] = sum_l Sy[l][j] work2[k][l][i] */
    for (int k = 0; k < Nqe; k++)
      for (int j = 0; j < Nqe; j++)
        for (int i = 0; i < Nqe; i++) {
          float value = 0.0f;
          for (int l = 0; l < Nqe; l++)
            value += SY(l, j) * work2[IDXE(k, l, i)];
          tmp[IDXE(j, k, i)] = value;
        }
    /* work2[k][i][j] = invL[k][j][i] sum_l Sz[l][k] tmp[j][l][i] */
    for (int k = 0; k < Nqe; k++)
      for (int j = 0; j < Nqe; j++)
        for (int i = 0; i < Nqe; i++) {
          const int v = i + j * Nqe + k * Nqe * Nqe;
          float value = 0.0f;
          for (int l = 0; l < Nqe; l++)
            value += SZ(l, k) * tmp[IDXE(j, l, i)];
          work2[IDXE(k, i, j)] = value * inv_L[v + (size_t)elem * Npe];
        }
    /* tmp[k][j][i] = sum_l SxT[l][i] work2[k][l][j],  SxT[l][i] = Sx[i][l] */
    for (int k = 0; k < Nqe; k++)
      for (int j = 0; j < Nqe; j++)
        for (int i = 0; i < Nqe; i++) {
          float value = 0.0f;
          for (int l = 0; l < Nqe; l++)
            value += SX(i, l) * work2[IDXE(k, l, j)];
          tmp[IDXE(k, j, i)] = value;
        }
    /* work2[j][k][i] = sum_l SyT[l][j] tmp[k][l][i] */
    for (int k = 0; k < Nqe; k++)
      for (int j = 0; j < Nqe; j++)
        for (int i = 0; i < Nqe; i++) {
          float value = 0.0f;
          for (int l = 0; l < Nqe; l++)
            value += SY(j, l) * tmp[IDXE(k, l, i)];
          work2[IDXE(j, k, i)] = value;
        }
    /* tmp[k][j][i] = sum_l SzT[l][k] work2[j][l][i] */
    for (int k = 0; k < Nqe; k++)
      for (int j = 0; j < Nqe; j++)
        for (int i = 0; i < Nqe; i++) {
          float value = 0.0f;
          for (int l = 0; l < Nqe; l++)
            value += SZ(k, l) * work2[IDXE(j, l, i)];
          if (!p_restrict)
            Su[(size_t)elem * Npe + IDXE(k, j, i)] = value;
          tmp[IDXE(k, j, i)] = value;
        }
    if (!p_restrict) {
      /* fusedFDM.c:166-205: only the six overlap planes of work2 are refreshed
       * from the solution; every other entry of work2 still holds step-5's
       * permuted intermediate and is written to u as is.                     */
      for (int a = 1; a < Nqe - 1; ++a)
        for (int b = 1; b < Nqe - 1; ++b) {
          work2[IDXE(0, a, b)] = tmp[IDXE(0, a, b)];
          work2[IDXE(Nqe - 1, a, b)] = tmp[IDXE(Nqe - 1, a, b)];
        }
      for (int a = 1; a < Nqe - 1; ++a)
        for (int b = 1; b < Nqe - 1; ++b)
          work2[IDXE(a, 0, b)] = tmp[IDXE(a, 0, b)];
      for (int a = 1; a < Nqe - 1; ++a)
        for (int b = 1; b < Nqe - 1; ++b)
          work2[IDXE(a, Nqe - 1, b)] = tmp[IDXE(a, Nqe - 1, b)];
      for (int a = 1; a < Nqe - 1; ++a)
        for (int b = 1; b < Nqe - 1; ++b)
          work2[IDXE(a, b, 0)] = tmp[IDXE(a, b, 0)];
      for (int a = 1; a < Nqe - 1; ++a)
        for (int b = 1; b < Nqe - 1; ++b)
          work2[IDXE(a, b, Nqe - 1)] = tmp[IDXE(a, b, Nqe - 1)];
      for (int n = 0; n < Npe; ++n)
        w1[n] = work2[n];
    } else {
      for (int k = 0; k < Nq; ++k)
        for (int j = 0; j < Nq; ++j)
          for (int i = 0; i < Nq; ++i) {
            const size_t idx = i + j * Nq + k * Nq * Nq + (size_t)elem * Nq * Nq * Nq;
            Su[idx] = tmp[IDXE(k + 1, j + 1, i + 1)] * wts[idx];
          }
    }
#undef SX
#undef SY
#undef SZ
  }
  free(Sx);
}

/* postFDM  kernels/elliptic/postFDM.c:14-153  (ASM only)                     */
void orc_post_fdm_f(dlong Nelements, const float *my_work1, const float *my_work2, float *Su, const float *wts,
                    int Nq)
{
  const int Nqe = Nq + 2, Npe = Nqe * Nqe * Nqe;
  float *work1 = (float *)malloc(sizeof(float) * 2 * (size_t)Npe);
  float *work2 = work1 + Npe;
  for (dlong elem = 0; elem < Nelements; ++elem) {
    for (int n = 0; n < Npe; ++n) {
      work1[n] = my_work2[n + (size_t)elem * Npe];
      work2[n] = my_work1[n + (size_t)elem * Npe];
    }
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b) {
        work1[IDXE(0, a, b)] = work1[IDXE(0, a, b)] - work2[IDXE(0, a, b)];
        work1[IDXE(Nqe - 1, a, b)] = work1[IDXE(Nqe - 1, a, b)] - work2[IDXE(Nqe - 1, a, b)];
      }
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b) {
        work1[IDXE(a, 0, b)] = work1[IDXE(a, 0, b)] - work2[IDXE(a, 0, b)];
        work1[IDXE(a, Nqe - 1, b)] = work1[IDXE(a, Nqe - 1, b)] - work2[IDXE(a, Nqe - 1, b)];
      }
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b) {
        work1[IDXE(a, b, 0)] = work1[IDXE(a, b, 0)] - work2[IDXE(a, b, 0)];
        work1[IDXE(a, b, Nqe - 1)] = work1[IDXE(a, b, Nqe - 1)] - work2[IDXE(a, b, Nqe - 1)];
      }
    /* the three directions must be folded one after another (postFDM.c:90-140):
     * plane 2 += plane 0 reads values the previous direction already updated  */
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        work1[IDXE(2, a, b)] = work1[IDXE(2, a, b)] + work1[IDXE(0, a, b)];
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        work1[IDXE(Nqe - 3, a, b)] = work1[IDXE(Nqe - 3, a, b)] + work1[IDXE(Nqe - 1, a, b)];
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        work1[IDXE(a, 2, b)] = work1[IDXE(a, 2, b)] + work1[IDXE(a, 0, b)];
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        work1[IDXE(a, Nqe - 3, b)] = work1[IDXE(a, Nqe - 3, b)] + work1[IDXE(a, Nqe - 1, b)];
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        work1[IDXE(a, b, 2)] = work1[IDXE(a, b, 2)] + work1[IDXE(a, b, 0)];
    for (int a = 1; a < Nqe - 1; ++a)
      for (int b = 1; b < Nqe - 1; ++b)
        work1[IDXE(a, b, Nqe - 3)] = work1[IDXE(a, b, Nqe - 3)] + work1[IDXE(a, b, Nqe - 1)];
    for (int k = 0; k < Nq; ++k)
      for (int j = 0; j < Nq; ++j)
        for (int i = 0; i < Nq; ++i) {
          const size_t idx = i + j * Nq + k * Nq * Nq + (size_t)elem * Nq * Nq * Nq;
          Su[idx] = work1[IDXE(k + 1, j + 1, i + 1)] * wts[idx];
        }
  }
  free(work1);
}
#undef IDXE

/* ------------------------------------------------------------------------- */
/* ellipticPreconCoarsenHex3D   kernels/elliptic/ellipticPreconCoarsenHex3D.c:26-100
 * pfloat in/out, dfloat accumulation.  R[NqC][NqF].                          */
void orc_coarsen_f(dlong Nelements, const float *R, const float *qf, float *qc, int NqF, int NqC)
{
  const int NpF = NqF * NqF * NqF, NpC = NqC * NqC * NqC;
  double *s_R = (double *)malloc(sizeof(double) * (NqC * NqF + (size_t)NqF * NqF * NqC + NqF * NqF + NqC * NqF));
  double *r_q = s_R + NqC * NqF; /* [NqF(j)][NqF(i)][NqC] */
  double *s_q = r_q + (size_t)NqF * NqF * NqC; /* [NqF(i)][NqF(j)] */
  double *s_Pq = s_q + NqF * NqF;              /* [NqC][NqF] */
  for (int t = 0; t < NqC * NqF; ++t)
    s_R[t] = R[t];
  for (dlong e = 0; e < Nelements; ++e) {
    for (int j = 0; j < NqF; ++j)
      for (int i = 0; i < NqF; ++i) {
        double *rq = r_q + ((size_t)j * NqF + i) * NqC;
        for (int k = 0; k < NqC; ++k)
          rq[k] = 0;
        for (int k = 0; k < NqF; ++k) {
          const double tmp = qf[i + j * NqF + k * NqF * NqF + (size_t)e * NpF];
          for (int m = 0; m < NqC; ++m)
            rq[m] += s_R[m * NqF + k] * tmp; /* s_RT[k][m] */
        }
      }
    for (int k = 0; k < NqC; ++k) {
      for (int j = 0; j < NqF; ++j)
        for (int i = 0; i < NqF; ++i)
          s_q[i * NqF + j] = r_q[((size_t)j * NqF + i) * NqC + k];
      for (int j = 0; j < NqC; ++j)
        for (int i = 0; i < NqF; ++i) {
          double res = 0;
          for (int m = 0; m < NqF; ++m)
            res += s_R[j * NqF + m] * s_q[i * NqF + m];
          s_Pq[j * NqF + i] = res;
        }
      for (int j = 0; j < NqC; ++j)
        for (int i = 0; i < NqC; ++i) {
          double res = 0;
          for (int m = 0; m < NqF; ++m)
            res += s_R[i * NqF + m] * s_Pq[j * NqF + m];
          qc[i + j * NqC + k * NqC * NqC + (size_t)e * NpC] = res;
        }
    }
  }
  free(s_R);
}

/* ellipticPreconProlongateHex3D  kernels/elliptic/ellipticPreconProlongateHex3D.c:26-110
 * qN += (R^T x R^T x R^T) qc, dfloat accumulation, pfloat I/O.               */
void orc_prolongate_f(dlong Nelements, const float *R, const float *qc, float *qN, int NqF, int NqC)
{
  const int NpF = NqF * NqF * NqF, NpC = NqC * NqC * NqC;
  double *s_R = (double *)malloc(sizeof(double) * (NqC * NqF + (size_t)NqC * NqC * NqF + NqC * NqC + NqF * NqC));
  double *r_q = s_R + NqC * NqF; /* [NqC*NqC (t)][NqF] */
  double *s_q = r_q + (size_t)NqC * NqC * NqF; /* [NqC][NqC] */
  double *s_Pq = s_q + NqC * NqC;              /* [NqF][NqC] */
  for (int t = 0; t < NqC * NqF; ++t)
    s_R[t] = R[t];
  for (dlong e = 0; e < Nelements; ++e) {
    for (int t = 0; t < NqC * NqC; ++t) {
      double *rq = r_q + (size_t)t * NqF;
      for (int k = 0; k < NqF; ++k)
        rq[k] = 0;
      for (int k = 0; k < NqC; ++k) {
        const double tmp = qc[t + k * NqC * NqC + (size_t)e * NpC];
        for (int m = 0; m < NqF; ++m)
          rq[m] += s_R[k * NqF + m] * tmp;
      }
    }
    for (int k = 0; k < NqF; ++k) {
      for (int t = 0; t < NqC * NqC; ++t)
        s_q[t] = r_q[(size_t)t * NqF + k]; /* s_q[tj][ti], t = ti + NqC tj */
      for (int t = 0; t < NqC * NqF; ++t) {
        const int ti = t % NqC, tj = t / NqC;
        double res = 0;
        for (int m = 0; m < NqC; ++m)
          res += s_R[m * NqF + tj] * s_q[m * NqC + ti];
        s_Pq[tj * NqC + ti] = res;
      }
      for (int j = 0; j < NqF; ++j)
        for (int i = 0; i < NqF; ++i) {
          double res = 0;
          for (int m = 0; m < NqC; ++m)
            res += s_R[m * NqF + i] * s_Pq[j * NqC + m];
          qN[i + j * NqF + k * NqF * NqF + (size_t)e * NpF] += res;
        }
    }
  }
  free(s_R);
}

/* ------------------------------------------------------------------------- */
/* GMRES kernels: gramSchmidtOrthogonalization.c, updatePGMRESSolution.c,
 * fusedResidualAndNorm.c                                                    */
double orc_gram_schmidt_d(dlong N, dlong offset, int Nfields, int gmresSize, const double *weights,
                          const double *y, const double *V, double *w)
{
  for (int j = 0; j < gmresSize; ++j) {
    const double yj = y[j];
    for (int fld = 0; fld < Nfields; fld++)
      for (dlong n = 0; n < N; ++n) {
        const double Vnj = V[n + fld * offset + (size_t)j * offset * Nfields];
        w[n + fld * offset] -= yj * Vnj;
      }
  }
  double sum = 0.0;
  for (int fld = 0; fld < Nfields; fld++)
    for (dlong n = 0; n < N; ++n) {
      const double weight = weights[n];
      const double w_curr = w[n + fld * offset];
      sum += w_curr * w_curr * weight;
    }
  return sum;
}
void orc_update_pgmres_solution_d(dlong N, dlong offset, int Nfields, int gmresSize, const double *y,
                                  const double *Z, double *x)
{
  for (int j = 0; j < gmresSize; ++j)
    for (int fld = 0; fld < Nfields; ++fld)
      for (int n = 0; n < N; ++n) {
        const double yj = y[j];
        const double Znj = Z[n + fld * offset + (size_t)j * offset * Nfields];
        x[n + fld * offset] += Znj * yj;
      }
}
double orc_fused_residual_and_norm_d(dlong N, dlong offset, int Nfields, const double *weights,
                                     const double *b_vec, const double *Ax, double *r)
{
  double rdotr = 0.0;
  for (int fld = 0; fld < Nfields; ++fld)
    for (int id = 0; id < N; ++id) {
      const double rnew = b_vec[id + fld * offset] - Ax[id + fld * offset];
      r[id + fld * offset] = rnew;
      rdotr += rnew * rnew * weights[id];
    }
  return rdotr;
}

/* ------------------------------------------------------------------------- */
/* geometricFactorsHex3D  kernels/mesh/geometricFactorsHex3D.okl:26-142 (one OKL
 * thread = one node).  Writes ggeo[E][7][Np] and the Jacobian.               */
void orc_geometric_factors_d(dlong Nelements, int Nq, const double *D, const double *gllw, const double *x,
                             const double *y, const double *z, double *ggeo, double *Jac)
{
  const int Np = Nq * Nq * Nq;
  for (dlong e = 0; e < Nelements; ++e) {
    const double *xe = x + (size_t)e * Np, *ye = y + (size_t)e * Np, *ze = z + (size_t)e * Np;
    for (int k = 0; k < Nq; ++k)
      for (int j = 0; j < Nq; ++j)
        for (int i = 0; i < Nq; ++i) {
          double xr = 0, yr = 0, zr = 0, xs = 0, ys = 0, zs = 0, xt = 0, yt = 0, zt = 0;
          for (int m = 0; m < Nq; ++m) {
            const double Dim = D[i * Nq + m], Djm = D[j * Nq + m], Dkm = D[k * Nq + m];
            const int r = k * Nq * Nq + j * Nq + m, s = k * Nq * Nq + m * Nq + i, t = m * Nq * Nq + j * Nq + i;
            xr += Dim * xe[r];
            xs += Djm * xe[s];
            xt += Dkm * xe[t];
            yr += Dim * ye[r];
            ys += Djm * ye[s];
            yt += Dkm * ye[t];
            zr += Dim * ze[r];
            zs += Djm * ze[s];
            zt += Dkm * ze[t];
          }
          const double J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt);
          const double Jinv = 1. / J;
          const double JW = J * gllw[i] * gllw[j] * gllw[k];
          const double rx = (ys * zt - zs * yt) * Jinv, ry = -(xs * zt - zs * xt) * Jinv,
                       rz = (xs * yt - ys * xt) * Jinv;
          const double sx = -(yr * zt - zr * yt) * Jinv, sy = (xr * zt - zr * xt) * Jinv,
                       sz = -(xr * yt - yr * xt) * Jinv;
          const double tx = (yr * zs - zr * ys) * Jinv, ty = -(xr * zs - zr * xs) * Jinv,
                       tz = (xr * ys - yr * xs) * Jinv;
          const int n = i + j * Nq + k * Nq * Nq;
          double *g = ggeo + (size_t)7 * Np * e + n;
          Jac[(size_t)e * Np + n] = J;
          g[0 * Np] = JW * (rx * rx + ry * ry + rz * rz);
          g[1 * Np] = JW * (rx * sx + ry * sy + rz * sz);
          g[4 * Np] = JW * (rx * tx + ry * ty + rz * tz);
          g[2 * Np] = JW * (sx * sx + sy * sy + sz * sz);
          g[3 * Np] = JW * (sx * tx + sy * ty + sz * tz);
          g[5 * Np] = JW * (tx * tx + ty * ty + tz * tz);
          g[6 * Np] = JW;
        }
  }
}

"""ctypes access to the checker libraries.

TEST INFRASTRUCTURE ONLY (see oracle/nrs_oracle.c header).

`Orc`  -> oracle/_oracle.so   (our C restatement, order N at run time)
`Ref*` -> oracle/_ref/*.so    (the reference's own serial kernels, compiled by
                               oracle/build_ref.py; OCCA serial ABI: scalars by
                               const reference, raw pointers)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")

c_int, c_dbl, c_flt = C.c_int, C.c_double, C.c_float


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _chk(a, dt):
    assert a.dtype == dt and a.flags["C_CONTIGUOUS"], (a.dtype, dt)
    return a


class Orc:
    def __init__(self, fast: bool = False):
        path = os.path.join(HERE, "_oracle_fast.so" if fast else "_oracle.so")
        if not os.path.exists(path):
            from . import build_ref
            build_ref.build_oracle()
        self.lib = C.CDLL(path)
        L = self.lib
        for s in ("d", "f"):
            getattr(L, "orc_weighted_inner_prod_many_" + s).restype = c_dbl if s == "d" else c_flt
            getattr(L, "orc_weighted_norm2_many_" + s).restype = c_dbl if s == "d" else c_flt
            getattr(L, "orc_sum_" + s).restype = c_dbl if s == "d" else c_flt
            getattr(L, "orc_inner_prod_" + s).restype = c_dbl if s == "d" else c_flt
        L.orc_gram_schmidt_d.restype = c_dbl
        L.orc_fused_residual_and_norm_d.restype = c_dbl

    @staticmethod
    def _suf(a):
        return "d" if a.dtype == np.float64 else "f"

    @staticmethod
    def _sc(a, v):
        return c_dbl(v) if a.dtype == np.float64 else c_flt(v)

    # --- operator
    def ax(self, N, element_list, ggeo, D, q, Aq, lambda0=None, lambda1=None, poisson=True):
        Nq = N + 1
        dt = q.dtype
        S = np.ascontiguousarray(D.T)
        lam0 = np.ones(1, dtype=dt) if lambda0 is None else lambda0
        lam1 = np.zeros(1, dtype=dt) if lambda1 is None else lambda1
        fn = getattr(self.lib, "orc_ax_" + self._suf(q))
        fn(c_int(len(element_list)), c_int(0), c_int(0), _p(_chk(element_list, np.int32)), _p(_chk(ggeo, dt)),
           _p(_chk(D, dt)), _p(S), _p(lam0), _p(lam1), _p(_chk(q, dt)), _p(_chk(Aq, dt)),
           c_int(Nq), c_int(1 if poisson else 0), c_int(0))
        return Aq

    def ax_block(self, N, element_list, ggeo, D, q, Aq, lambda0, lambda1, offset, loffset, lambda_field=False,
                 Nfields=3):
        """ellipticBlockPartialAxCoeffHex3D.c:2-152: the scalar Helmholtz operator field by field (q, Aq `offset`
        apart, coefficients `loffset` apart); the per-node arithmetic of the reference's block kernel is the scalar
        kernel's, statement for statement."""
        Nq = N + 1
        dt = q.dtype
        S = np.ascontiguousarray(D.T)
        fn = getattr(self.lib, "orc_ax_" + self._suf(q))
        for f in range(Nfields):
            fn(c_int(len(element_list)), c_int(0), c_int(0), _p(_chk(element_list, np.int32)), _p(_chk(ggeo, dt)),
               _p(_chk(D, dt)), _p(S), _p(lambda0[f * loffset:]), _p(lambda1[f * loffset:]),
               _p(q[f * offset:]), _p(Aq[f * offset:]), c_int(Nq), c_int(0), c_int(1 if lambda_field else 0))
        return Aq

    def ax_stress(self, N, element_list, vgeo, D, q, Aq, lambda0, lambda1, offset, loffset, lambda_field=False):
        """ellipticStressPartialAxCoeffHex3D.c:1-169: three coupled fields, stress form; vgeo = 12 planes per
        element (rx..tz, J, JW, 1/JW)."""
        dt = q.dtype
        fn = getattr(self.lib, "orc_ax_stress_" + self._suf(q))
        fn(c_int(len(element_list)), c_int(offset), c_int(loffset), _p(_chk(element_list, np.int32)),
           _p(_chk(vgeo, dt)), _p(_chk(D, dt)), _p(_chk(lambda0, dt)), _p(_chk(lambda1, dt)), _p(_chk(q, dt)),
           _p(_chk(Aq, dt)), c_int(N + 1), c_int(1 if lambda_field else 0))
        return Aq

    def mask(self, mask_ids, q):
        getattr(self.lib, "orc_mask_" + self._suf(q))(c_int(len(mask_ids)), _p(_chk(mask_ids, np.int32)), _p(q))

    def gs_add(self, ogs, q, k=1, stride=0):
        getattr(self.lib, "orc_gs_add_" + self._suf(q))(c_int(ogs.Ngather), _p(ogs.offsets), _p(ogs.gather_ids),
                                                        c_int(k), c_int(stride), _p(q))

    def gs_min_i(self, ogs, q):
        self.lib.orc_gs_min_i(c_int(ogs.Ngather), _p(ogs.offsets), _p(ogs.gather_ids), _p(_chk(q, np.int32)))

    # --- Krylov helpers
    def update_pcg(self, N, inv_degree, Ap, alpha, r):
        out = np.zeros(1)
        self.lib.orc_update_pcg_d(c_int(N), c_int(0), c_int(1), _p(inv_degree), _p(Ap), c_dbl(alpha), _p(r), _p(out))
        return float(out[0])

    def axpby(self, N, a, x, b, y):
        getattr(self.lib, "orc_axpby_many_" + self._suf(y))(c_int(N), c_int(1), c_int(0), self._sc(y, a), _p(x),
                                                            self._sc(y, b), _p(y))

    def axmyz(self, N, a, x, y, z):
        getattr(self.lib, "orc_axmyz_" + self._suf(z))(c_int(N), self._sc(z, a), _p(x), _p(y), _p(z))

    def axmy(self, N, a, x, y):
        getattr(self.lib, "orc_axmy_" + self._suf(y))(c_int(N), self._sc(y, a), _p(x), _p(y))

    def weighted_inner_prod(self, N, w, x, y):
        return float(getattr(self.lib, "orc_weighted_inner_prod_many_" + self._suf(x))(
            c_int(N), c_int(1), c_int(0), _p(w), _p(x), _p(y)))

    def weighted_norm2_sq(self, N, w, x):
        return float(getattr(self.lib, "orc_weighted_norm2_many_" + self._suf(x))(
            c_int(N), c_int(1), c_int(0), _p(w), _p(x)))

    def weighted_inner_prod_multi(self, N, NVec, offset, w, x, y):
        out = np.zeros(NVec)
        self.lib.orc_weighted_inner_prod_multi_d(c_int(N), c_int(NVec), c_int(offset), _p(w), _p(x), _p(y), _p(out))
        return out

    def sum(self, N, x):
        return float(getattr(self.lib, "orc_sum_" + self._suf(x))(c_int(N), _p(x)))

    def copy_d2f(self, x, y):
        self.lib.orc_copy_d2f(c_int(x.size), _p(_chk(x, np.float64)), _p(_chk(y, np.float32)))

    def copy_f2d(self, x, y):
        self.lib.orc_copy_f2d(c_int(x.size), _p(_chk(x, np.float32)), _p(_chk(y, np.float64)))

    def gram_schmidt(self, N, offset, gmres_size, w, y, V, wv):
        return float(self.lib.orc_gram_schmidt_d(c_int(N), c_int(offset), c_int(1), c_int(gmres_size), _p(w),
                                                 _p(y), _p(V), _p(wv)))

    def update_pgmres_solution(self, N, offset, gmres_size, y, Z, x):
        self.lib.orc_update_pgmres_solution_d(c_int(N), c_int(offset), c_int(1), c_int(gmres_size), _p(y), _p(Z), _p(x))

    def fused_residual_and_norm(self, N, w, b, Ax, r):
        return float(self.lib.orc_fused_residual_and_norm_d(c_int(N), c_int(0), c_int(1), _p(w), _p(b), _p(Ax), _p(r)))

    # --- multigrid
    def update_chebyshev(self, N, dCoeff, rCoeff, SAd, d, r, x):
        self.lib.orc_update_chebyshev_f(c_int(N), c_flt(dCoeff), c_flt(rCoeff), _p(SAd), _p(d), _p(r), _p(x))

    def update_fourth_chebyshev(self, N, beta, Ad, d, r, x):
        self.lib.orc_update_fourth_chebyshev_f(c_int(N), c_flt(beta), _p(Ad), _p(d), _p(r), _p(x))

    def pre_fdm(self, E, N, u, work1):
        self.lib.orc_pre_fdm_f(c_int(E), _p(_chk(u, np.float32)), _p(_chk(work1, np.float32)), c_int(N + 1))

    def fused_fdm(self, E, N, Su, Sx, Sy, Sz, invL, wts, u, restrict):
        el = np.arange(E, dtype=np.int32)
        self.lib.orc_fused_fdm_f(c_int(E), _p(el), _p(Su), _p(Sx), _p(Sy), _p(Sz), _p(invL), _p(wts), _p(u),
                                 c_int(N + 1), c_int(restrict))

    def post_fdm(self, E, N, work1, work2, Su, wts):
        self.lib.orc_post_fdm_f(c_int(E), _p(work1), _p(work2), _p(Su), _p(wts), c_int(N + 1))

    def coarsen(self, E, Nf, Nc, R, qf, qc):
        self.lib.orc_coarsen_f(c_int(E), _p(_chk(R, np.float32)), _p(_chk(qf, np.float32)), _p(_chk(qc, np.float32)),
                               c_int(Nf + 1), c_int(Nc + 1))

    def prolongate(self, E, Nf, Nc, R, qc, qN):
        self.lib.orc_prolongate_f(c_int(E), _p(_chk(R, np.float32)), _p(_chk(qc, np.float32)),
                                  _p(_chk(qN, np.float32)), c_int(Nf + 1), c_int(Nc + 1))

    def geometric_factors(self, E, N, D, gllw, x, y, z):
        Np = (N + 1) ** 3
        ggeo = np.empty((E, 7, Np))
        J = np.empty((E, Np))
        self.lib.orc_geometric_factors_d(c_int(E), c_int(N + 1), _p(D), _p(gllw), _p(x), _p(y), _p(z), _p(ggeo), _p(J))
        return ggeo, J


# ----------------------------------------------------------------------------- reference
def ref_available(name: str) -> bool:
    return os.path.exists(os.path.join(REFDIR, name + ".so"))


def _ref(name):
    return C.CDLL(os.path.join(REFDIR, name + ".so"))


def _r(v, t=c_int):
    return C.byref(t(v))


class RefAx:
    """ellipticPartialAxCoeffHex3D_v0 compiled from the reference."""

    def __init__(self, N, prec="d", poisson=True, fast=False):
        self.N, self.prec = N, prec
        self.lib = _ref("ax_%s_N%d_%s%s" % (prec, N, "poisson" if poisson else "helmholtz", "_fast" if fast else ""))

    def __call__(self, element_list, ggeo, D, q, Aq, lambda0=None, lambda1=None):
        dt = q.dtype
        S = np.ascontiguousarray(D.T)
        lam0 = np.ones(1, dtype=dt) if lambda0 is None else lambda0
        lam1 = np.zeros(1, dtype=dt) if lambda1 is None else lambda1
        self.lib.ellipticPartialAxCoeffHex3D_v0(_r(len(element_list)), _r(0), _r(0), _p(element_list), _p(ggeo),
                                                _p(D), _p(S), _p(lam0), _p(lam1), _p(q), _p(Aq))
        return Aq


class RefAxBlock:
    """ellipticBlockPartialAxCoeffHex3D_v0 compiled from the reference (three fields)."""

    def __init__(self, N, lambda_field):
        self.lib = _ref("axblock_d_N%d_lambda%d" % (N, 1 if lambda_field else 0))

    def __call__(self, element_list, ggeo, D, q, Aq, lambda0, lambda1, offset, loffset):
        S = np.ascontiguousarray(D.T)
        self.lib.ellipticBlockPartialAxCoeffHex3D_v0(_r(len(element_list)), _r(offset), _r(loffset), _p(element_list),
                                                     _p(ggeo), _p(D), _p(S), _p(lambda0), _p(lambda1), _p(q), _p(Aq))
        return Aq


class RefAxStress:
    """ellipticStressPartialAxCoeffHex3D_v0 compiled from the reference (three coupled fields)."""

    def __init__(self, N, lambda_field):
        self.lib = _ref("axstress_d_N%d_lambda%d" % (N, 1 if lambda_field else 0))

    def __call__(self, element_list, vgeo, D, q, Aq, lambda0, lambda1, offset, loffset):
        S = np.ascontiguousarray(D.T)
        self.lib.ellipticStressPartialAxCoeffHex3D_v0(_r(len(element_list)), _r(offset), _r(loffset), _p(element_list),
                                                      _p(vgeo), _p(D), _p(S), _p(lambda0), _p(lambda1), _p(q), _p(Aq))
        return Aq


class RefFdm:
    def __init__(self, N, restrict, fast=False):
        self.N, self.restrict = N, restrict
        self.lib = _ref("fdm_N%d_r%d%s" % (N, restrict, "_fast" if fast else ""))

    def pre(self, E, u, work1):
        self.lib.preFDM(_r(E), _p(u), _p(work1))

    def fused(self, E, Su, Sx, Sy, Sz, invL, wts, u):
        el = np.arange(E, dtype=np.int32)
        if self.restrict:
            self.lib.fusedFDM(_r(E), _p(el), _p(Su), _p(Sx), _p(Sy), _p(Sz), _p(invL), _p(wts), _p(u))
        else:
            self.lib.fusedFDM(_r(E), _p(el), _p(Su), _p(Sx), _p(Sy), _p(Sz), _p(invL), _p(u))

    def post(self, E, work1, work2, Su, wts):
        self.lib.postFDM(_r(E), _p(work1), _p(work2), _p(Su), _p(wts))


class RefTransfer:
    def __init__(self, Nf, Nc):
        self.lib = _ref("transfer_Nf%d_Nc%d" % (Nf, Nc))

    def coarsen(self, E, R, qf, qc):
        self.lib.ellipticPreconCoarsenHex3D(_r(E), _p(R), _p(qf), _p(qc))

    def prolongate(self, E, R, qc, qN):
        self.lib.ellipticPreconProlongateHex3D(_r(E), _p(R), _p(qc), _p(qN))


class RefLinAlg:
    def __init__(self, prec="d", fast=False):
        self.lib = _ref("linalg_%s%s" % (prec, "_fast" if fast else ""))
        self.t = c_dbl if prec == "d" else c_flt
        self.dt = np.float64 if prec == "d" else np.float32

    def axpby_many(self, N, a, x, b, y):
        self.lib.axpbyMany(_r(N), _r(1), _r(0), _r(a, self.t), _p(x), _r(b, self.t), _p(y))

    def axmyz(self, N, a, x, y, z):
        self.lib.axmyz(_r(N), _r(a, self.t), _p(x), _p(y), _p(z))

    def axmy(self, N, a, x, y):
        self.lib.axmy(_r(N), _r(a, self.t), _p(x), _p(y))

    def weighted_inner_prod_many(self, N, w, x, y):
        out = np.zeros(1, dtype=self.dt)
        self.lib.weightedInnerProdMany(_r(1), _r(N), _r(1), _r(0), _p(w), _p(x), _p(y), _p(out))
        return float(out[0])

    def weighted_norm2_many(self, N, w, x):
        out = np.zeros(1, dtype=self.dt)
        self.lib.weightedNorm2Many(_r(1), _r(N), _r(1), _r(0), _p(w), _p(x), _p(out))
        return float(out[0])

    def update_pcg(self, N, inv_degree, Ap, alpha, r):
        out = np.zeros(1, dtype=self.dt)
        self.lib.ellipticBlockUpdatePCG(_r(N), _r(0), _p(inv_degree), _p(Ap), _r(alpha, self.t), _p(r), _p(out))
        return float(out[0])

    def gram_schmidt(self, N, offset, gmres_size, w, y, V, wv):
        out = np.zeros(1, dtype=self.dt)
        self.lib.gramSchmidtOrthogonalization(_r(1), _r(N), _r(offset), _r(gmres_size), _p(w), _p(y), _p(V), _p(wv),
                                              _p(out))
        return float(out[0])

    def update_pgmres_solution(self, N, offset, gmres_size, y, Z, x):
        self.lib.updatePGMRESSolution(_r(N), _r(offset), _r(gmres_size), _p(y), _p(Z), _p(x))

    def fused_residual_and_norm(self, N, w, b, Ax, r):
        out = np.zeros(1, dtype=self.dt)
        self.lib.fusedResidualAndNorm(_r(1), _r(N), _r(0), _p(w), _p(b), _p(Ax), _p(r), _p(out))
        return float(out[0])

    def copy_d2f(self, x, y):
        self.lib.copyDfloatToPfloat(_r(x.size), _p(x), _p(y))

    def copy_f2d(self, x, y):
        self.lib.copyPfloatToDfloat(_r(x.size), _p(x), _p(y))


def build_diagonal(N, E, ggeo, D, lambda0, lambda1, poisson=True, lambda_field=False, Nfields=1, offset=None,
                   loffset=0):
    """ellipticBlockBuildDiagonalHex3D.okl:26-118 (OKL only): Aq[id + l*offset] = element diagonal of
    D^T (lambda0 G) D (+ lambda1 GwJ).  numpy restatement of the thread body; the accumulation runs in the dtype of
    ggeo like the kernel's dfloat / pfloat."""
    Nq = N + 1
    Np = Nq ** 3
    dt = ggeo.dtype
    G = ggeo.reshape(E, 7, Nq, Nq, Nq)
    Dd = np.asarray(D, dtype=dt)
    D2 = Dd * Dd
    dd = np.diag(Dd).astype(dt)
    offset = E * Np if offset is None else offset
    out = np.zeros(max(Nfields * offset, E * Np), dtype=dt)
    for l in range(Nfields):
        if lambda_field:
            l0 = lambda0[l * loffset:l * loffset + E * Np].reshape(E, Nq, Nq, Nq).astype(dt)
        else:
            l0 = np.full((E, Nq, Nq, Nq), lambda0[0], dtype=dt)
        r = np.zeros((E, Nq, Nq, Nq), dtype=dt)
        r += dt.type(2) * G[:, 1] * l0 * dd[None, None, None, :] * dd[None, None, :, None]
        r += dt.type(2) * G[:, 4] * l0 * dd[None, None, None, :] * dd[None, :, None, None]
        r += dt.type(2) * G[:, 3] * l0 * dd[None, None, :, None] * dd[None, :, None, None]
        r += np.einsum("ekjm,mi->ekji", G[:, 0] * l0, D2).astype(dt)
        r += np.einsum("ekmi,mj->ekji", G[:, 2] * l0, D2).astype(dt)
        r += np.einsum("emji,mk->ekji", G[:, 5] * l0, D2).astype(dt)
        if not poisson:
            if lambda_field:
                l1 = lambda1[l * loffset:l * loffset + E * Np].reshape(E, Nq, Nq, Nq).astype(dt)
            else:
                l1 = dt.type(lambda1[0])
            r += G[:, 6] * l1
        out[l * offset:l * offset + E * Np] = r.reshape(-1)
    return out


"""Kernel-level operators: thin Python mirrors of the reference kernel calls
(`elliptic->AxKernel(...)`, `platform->linAlg->...`, `oogs::startFinish(...)`), each forwarding
to the C ABI.  Arguments are DeviceBuffers / raw device addresses; nothing is computed in Python.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lib import DeviceBuffer, call, vp, i32, f64, f32


def _prec(dtype) -> int:
    return np.dtype(dtype).itemsize


def ellipticPartialAxCoeffHex3D(N, element_list, ggeo, D_host, q, Aq, *, Nelements=None, lambda0=None, lambda1=None,
                                poisson=True, lambda_field=False, variant=-1, stream=None, dtype=None):
    """AxKernel(NelementsList, fieldOffset, loffset, o_elementList, o_ggeo, o_D, o_DT, o_lambda0,
    o_lambda1, o_q, o_Aq)  -- ellipticOperator.cpp:83-93."""
    dtype = np.dtype(dtype or q.dtype)
    D_host = np.ascontiguousarray(D_host, dtype=dtype)
    ne = element_list.size if Nelements is None else Nelements
    call("nrsb_ellipticPartialAxCoeffHex3D", C.c_int(N + 1), C.c_int(_prec(dtype)), C.c_int(variant), i32(ne),
         i32(0), i32(0), vp(element_list), vp(ggeo), vp(D_host), vp(lambda0), vp(lambda1),
         C.c_int(1 if poisson else 0), C.c_int(1 if lambda_field else 0), vp(q), vp(Aq), vp(stream))


def mask(mask_ids, q, *, stream=None):
    call("nrsb_mask", C.c_int(_prec(q.dtype)), i32(mask_ids.size), vp(mask_ids), vp(q), vp(stream))


def gatherScatterMany_add(Ngather, starts, ids, q, *, k=1, stride=0, stream=None):
    call("nrsb_gatherScatterMany_add", C.c_int(_prec(q.dtype)), i32(Ngather), C.c_int(k), i32(stride), vp(starts),
         vp(ids), vp(q), vp(stream))


def fill(N, alpha, a, *, stream=None):
    call("nrsb_fill", C.c_int(_prec(a.dtype)), i32(N), f64(alpha), vp(a), vp(stream))


def axpby(N, alpha, x, beta, y, *, stream=None):
    call("nrsb_axpby", C.c_int(_prec(y.dtype)), i32(N), f64(alpha), vp(x), f64(beta), vp(y), vp(stream))


def axpbyMany(N, Nfields, offset, alpha, x, beta, y, *, stream=None):
    call("nrsb_axpbyMany", C.c_int(_prec(y.dtype)), i32(N), C.c_int(Nfields), i32(offset), f64(alpha), vp(x),
         f64(beta), vp(y), vp(stream))


def axmyz(N, alpha, x, y, z, *, stream=None):
    call("nrsb_axmyz", C.c_int(_prec(z.dtype)), i32(N), f64(alpha), vp(x), vp(y), vp(z), vp(stream))


def scale(N, alpha, x, *, stream=None):
    call("nrsb_scale", C.c_int(_prec(x.dtype)), i32(N), f64(alpha), vp(x), vp(stream))


def copyDfloatToPfloat(N, x, y, *, stream=None):
    call("nrsb_copyDfloatToPfloat", i32(N), vp(x), vp(y), vp(stream))


def copyPfloatToDfloat(N, x, y, *, stream=None):
    call("nrsb_copyPfloatToDfloat", i32(N), vp(x), vp(y), vp(stream))


def weightedInnerProdMany(N, Nfields, offset, w, x, y, *, stream=None) -> float:
    out = f64(0)
    call("nrsb_weightedInnerProdMany", i32(N), C.c_int(Nfields), i32(offset), vp(w), vp(x), vp(y), C.byref(out),
         vp(stream))
    return out.value


def weightedNorm2Many(N, Nfields, offset, w, x, *, stream=None) -> float:
    """returns sum w x^2 (the reference takes the sqrt in linAlg.cpp)"""
    out = f64(0)
    call("nrsb_weightedNorm2Many", i32(N), C.c_int(Nfields), i32(offset), vp(w), vp(x), C.byref(out), vp(stream))
    return out.value


def weightedInnerProdMulti(N, NVec, offset, w, x, y, *, stream=None) -> np.ndarray:
    out = np.zeros(NVec)
    call("nrsb_weightedInnerProdMulti", i32(N), C.c_int(NVec), i32(offset), vp(w), vp(x), vp(y), vp(out), vp(stream))
    return out


def sum(N, x, *, stream=None) -> float:  # noqa: A001  (linAlg_t::sum)
    out = f64(0)
    call("nrsb_sum", C.c_int(_prec(x.dtype)), i32(N), vp(x), C.byref(out), vp(stream))
    return out.value


def ellipticBlockUpdatePCG(N, inv_degree, Ap, alpha, r, *, p=None, x=None, stream=None) -> float:
    out = f64(0)
    call("nrsb_ellipticBlockUpdatePCG", i32(N), i32(0), vp(inv_degree), vp(Ap), f64(alpha), vp(r), vp(p), vp(x),
         C.byref(out), vp(stream))
    return out.value


def updateChebyshev(N, dCoeff, rCoeff, SAd, d, r, x, *, stream=None):
    call("nrsb_updateChebyshev", i32(N), f32(dCoeff), f32(rCoeff), vp(SAd), vp(d), vp(r), vp(x), vp(stream))


def updateFourthKindChebyshev(N, beta, Ad, d, r, x, *, stream=None):
    call("nrsb_updateFourthKindChebyshev", i32(N), f32(beta), vp(Ad), vp(d), vp(r), vp(x), vp(stream))


def gramSchmidtOrthogonalization(N, offset, gmres_size, weights, y, V, w, *, stream=None) -> float:
    out = f64(0)
    call("nrsb_gramSchmidtOrthogonalization", i32(N), i32(offset), C.c_int(gmres_size), vp(weights), vp(y), vp(V),
         vp(w), C.byref(out), vp(stream))
    return out.value


def updatePGMRESSolution(N, offset, gmres_size, y, Z, x, *, stream=None):
    call("nrsb_updatePGMRESSolution", i32(N), i32(offset), C.c_int(gmres_size), vp(y), vp(Z), vp(x), vp(stream))


def fusedResidualAndNorm(N, weights, b, Ax, r, *, stream=None) -> float:
    out = f64(0)
    call("nrsb_fusedResidualAndNorm", i32(N), i32(0), vp(weights), vp(b), vp(Ax), vp(r), C.byref(out), vp(stream))
    return out.value


def preFDM(N, Nelements, u, work1, *, stream=None):
    call("nrsb_preFDM", C.c_int(N + 1), i32(Nelements), vp(u), vp(work1), vp(stream))


def fusedFDM(N, restrict, Nelements, element_list, Su, Sx, Sy, Sz, invL, wts, u, *, stream=None):
    call("nrsb_fusedFDM", C.c_int(N + 1), C.c_int(restrict), i32(Nelements), vp(element_list), vp(Su), vp(Sx),
         vp(Sy), vp(Sz), vp(invL), vp(wts), vp(u), vp(stream))


def postFDM(N, Nelements, work1, work2, Su, wts, *, stream=None):
    call("nrsb_postFDM", C.c_int(N + 1), i32(Nelements), vp(work1), vp(work2), vp(Su), vp(wts), vp(stream))


def ellipticPreconCoarsenHex3D(Nf, Nc, Nelements, R_host, qf, qc, *, stream=None):
    R_host = np.ascontiguousarray(R_host, dtype=np.float32)
    call("nrsb_ellipticPreconCoarsenHex3D", C.c_int(Nf + 1), C.c_int(Nc + 1), i32(Nelements), vp(R_host), vp(qf),
         vp(qc), vp(stream))


def ellipticPreconProlongateHex3D(Nf, Nc, Nelements, R_host, qc, qN, *, stream=None):
    R_host = np.ascontiguousarray(R_host, dtype=np.float32)
    call("nrsb_ellipticPreconProlongateHex3D", C.c_int(Nf + 1), C.c_int(Nc + 1), i32(Nelements), vp(R_host), vp(qc),
         vp(qN), vp(stream))


def geometricFactorsHex3D(N, Nelements, D_host, gllw_host, x, y, z, ggeo, jac=None, *, stream=None):
    call("nrsb_geometricFactorsHex3D", C.c_int(N + 1), i32(Nelements), vp(np.ascontiguousarray(D_host)),
         vp(np.ascontiguousarray(gllw_host)), vp(x), vp(y), vp(z), vp(ggeo), vp(jac), vp(stream))


class Ogs:
    """ogs_t handle (ogsSetup, ogs.hpp:145-180) + single-rank oogs::startFinish."""

    def __init__(self, ids: np.ndarray, topo=None):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        self._h = C.c_void_p()
        self._ids = ids
        call("nrsb_ogs_setup", i32(ids.size), vp(ids), vp(None) if topo is None else C.byref(topo), C.byref(self._h))
        s = (C.c_int64 * 9)()
        call("nrsb_ogs_sizes", self._h, s)
        (self.N, self.Nlocal, self.NlocalGather, self.Nhalo, self.NhaloGather, self.nPairs, self.nQuads, self.nOcts,
         self.nGen) = [int(v) for v in s]

    def local_maps(self):
        off = np.zeros(self.NlocalGather + 1, dtype=np.int32)
        ids = np.zeros(max(self.Nlocal, 1), dtype=np.int32)
        call("nrsb_ogs_get_local_maps", self._h, vp(off), vp(ids))
        return off, ids[:self.Nlocal]

    def inv_degree(self):
        out = np.zeros(self.N)
        call("nrsb_ogs_get_inv_degree", self._h, vp(out))
        return out

    def inv_degree_device(self):
        d, f = C.c_void_p(), C.c_void_p()
        call("nrsb_ogs_inv_degree_device", self._h, C.byref(d), C.byref(f))
        return d.value, f.value

    def gather_scatter(self, v, *, k=1, stride=0, mask_ids=None, stream=None, dtype=None):
        dtype = np.dtype(dtype or v.dtype)
        nm = 0 if mask_ids is None else mask_ids.size
        call("nrsb_ogs_gather_scatter", self._h, C.c_int(dtype.itemsize), C.c_int(k), i32(stride), i32(nm),
             vp(mask_ids), vp(v), vp(stream))

    def destroy(self):
        if self._h:
            call("nrsb_ogs_destroy", self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def ellipticBlockBuildDiagonalHex3D(N, Nelements, ggeo, D_host, lambda0, lambda1, Aq, *, Nfields=1, offset=0,
                                    loffset=0, poisson=True, lambda_field=False, dtype=np.float64, stream=None):
    prec = 8 if np.dtype(dtype) == np.float64 else 4
    D_host = np.ascontiguousarray(D_host, dtype=dtype)
    call("nrsb_ellipticBlockBuildDiagonalHex3D", C.c_int(N + 1), C.c_int(prec), i32(Nelements), C.c_int(Nfields),
         i32(offset), i32(loffset), vp(ggeo), vp(D_host), vp(lambda0), vp(lambda1), C.c_int(1 if poisson else 0),
         C.c_int(1 if lambda_field else 0), vp(Aq), vp(stream))


def linalg_many(name, prec, *args, stream=None):
    """nrsb_scaleMany / nrsb_axmyMany / nrsb_axmyzMany / nrsb_adyMany / nrsb_axpbyzMany / nrsb_add / nrsb_axmy /
    nrsb_axdy / nrsb_abs with ctypes-converted arguments (ints -> int32, floats -> double, buffers -> void*)."""
    conv = [C.c_int(prec)]
    for a in args:
        if isinstance(a, (int, np.integer)):
            conv.append(i32(int(a)))
        elif isinstance(a, float):
            conv.append(C.c_double(a))
        else:
            conv.append(vp(a))
    conv.append(vp(stream))
    call("nrsb_" + name, *conv)


def ellipticStressPartialAxCoeffHex3D(N, Nelements, offset, loffset, element_list, vgeo, D_host, lambda0, lambda1, q, Aq,
                                      *, lambda_field=False, dtype=np.float64, stream=None):
    """AxKernel of a stress-form block solver (ellipticStressPartialAxCoeffHex3D.okl); vgeo = 12 planes per element."""
    prec = 8 if np.dtype(dtype) == np.float64 else 4
    D_host = np.ascontiguousarray(D_host, dtype=dtype)
    call("nrsb_ellipticStressPartialAxCoeffHex3D", C.c_int(N + 1), C.c_int(prec), i32(Nelements), i32(offset),
         i32(loffset), vp(element_list), vp(vgeo), vp(D_host), vp(lambda0), vp(lambda1),
         C.c_int(1 if lambda_field else 0), vp(q), vp(Aq), vp(stream))


def ellipticBlockPartialAxCoeffHex3D(N, Nelements, offset, loffset, element_list, ggeo, D_host, lambda0, lambda1, q, Aq,
                                     *, lambda_field=False, dtype=np.float64, stream=None):
    prec = 8 if np.dtype(dtype) == np.float64 else 4
    D_host = np.ascontiguousarray(D_host, dtype=dtype)
    call("nrsb_ellipticBlockPartialAxCoeffHex3D", C.c_int(N + 1), C.c_int(prec), i32(Nelements), i32(offset),
         i32(loffset), vp(element_list), vp(ggeo), vp(D_host), vp(lambda0), vp(lambda1),
         C.c_int(1 if lambda_field else 0), vp(q), vp(Aq), vp(stream))


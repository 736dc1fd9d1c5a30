"""CPU restatement of the reference's HOST control flow for the pressure solve, driving only the
oracle kernels (oracle/nrs_oracle.c) -- the stand-in for "nekRS on the SERIAL backend", which
cannot be built here (needs MPI + Fortran).

TEST INFRASTRUCTURE ONLY (see oracle/nrs_oracle.c header).

Follows, function by function: ellipticSetup.cpp:116-327, ellipticOgs.cpp:4-134,
ellipticOperator.cpp:31-172, ellipticSolve.cpp:32-190, PCG.cpp:33-203, PGMRES.cpp:31-340,
ellipticPreconditioner.cpp:33-84, MGSolver.cpp:150-193, ellipticMultiGridLevel.cpp:32-277,
ellipticMultiGridLevelSetup.cpp:109-453, ellipticMultiGridSchwarz.cpp:58-1156,
ellipticSolutionProjection.cpp:44-288, determineMGLevels.cpp:58-95.

Single rank.  A multi-GPU run of the product is checked against this driver on the WHOLE mesh.
Deviations from the reference, shared with the product and stated in DESIGN.md:
  * Arnoldi start vector = splitmix64 hash of the global node id (reference: std::random_device);
  * coarse solve = Jacobi-PCG on the assembled N=1 operator (reference: hypre BoomerAMG).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from . import sem
from .kernels import Orc
from .optimal_coeffs import optimal_coeffs

f32 = np.float32


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def id_uniform(ids: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return (splitmix64(ids) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def compare(opts, key, token):
    return token in opts.get(key, "")


class OMesh:
    def __init__(self, orc: Orc, N, E, x, y, z, global_ids, EToB):
        self.N, self.Nq, self.Np, self.E = N, N + 1, (N + 1) ** 3, E
        self.Nlocal = E * self.Np
        self.gllz, self.gllw = sem.jacobi_gll(N)
        self.D = sem.dmatrix_1d(self.gllz)
        self.x, self.y, self.z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
        self.global_ids = np.asarray(global_ids, dtype=np.int64)
        self.EToB = np.asarray(EToB, dtype=np.int32)
        self.ggeo, J = orc.geometric_factors(E, N, self.D, self.gllw, self.x, self.y, self.z)
        self.ggeo_f = self.ggeo.astype(f32)
        self.D_f = self.D.astype(f32)
        self.volume = float(self.ggeo[:, 6].sum())
        self.ogs = sem.Ogs(self.global_ids)
        self.element_list = np.arange(E, dtype=np.int32)


class OElliptic:
    """elliptic_t for one mesh (solver or MG level)."""

    def __init__(self, orc: Orc, mesh: OMesh, options: dict, poisson=True, lambda0=1.0, lambda1=0.0):
        self.orc, self.mesh, self.options = orc, mesh, options
        # constant coefficients: A = lambda0 * stiffness [+ lambda1 * mass]  (p_poisson, ellipticOperator.cpp:83-93)
        self.poisson, self.lambda0, self.lambda1 = bool(poisson), float(lambda0), float(lambda1)
        self.mask_ids, _ = sem.dirichlet_mask_ids(mesh.N, mesh.E, mesh.EToB, mesh.ogs, orc)
        ids = mesh.global_ids.copy()
        ids[self.mask_ids] = 0
        self.ogs = sem.Ogs(ids)
        self.inv_degree = self.ogs.inv_degree
        self.inv_degree_f = self.inv_degree.astype(f32)
        etob = mesh.EToB
        # the null space only exists for the pure Poisson operator (ellipticSetup.cpp:182-204)
        self.allNeumann = int(self.poisson and not np.any((etob > 0) & (etob != 4)))

    def ax(self, q, Aq):
        m = self.mesh
        dt = q.dtype.type
        g, D = (m.ggeo, m.D) if q.dtype == np.float64 else (m.ggeo_f, m.D_f)
        if self.poisson and self.lambda0 == 1.0:
            self.orc.ax(m.N, m.element_list, g, D, q, Aq)
        else:
            self.orc.ax(m.N, m.element_list, g, D, q, Aq, np.array([self.lambda0], dtype=dt),
                        np.array([self.lambda1], dtype=dt), poisson=self.poisson)

    def apply_mask(self, v):
        self.orc.mask(self.mask_ids, v)

    def gs(self, v):
        self.orc.gs_add(self.ogs, v)

    def operator(self, q, Aq, masked=True):
        self.ax(q, Aq)
        if masked:
            self.apply_mask(Aq)
        self.gs(Aq)

    def build_inv_diag(self, dtype):
        """ellipticBlockBuildDiagonalHex3D.okl + gs + 1/x (ellipticUpdateJacobi.cpp:32-85)."""
        m = self.mesh
        Nq = m.Nq
        G = (m.ggeo if dtype == np.float64 else m.ggeo_f.astype(np.float64)).reshape(m.E, 7, Nq, Nq, Nq)
        D = m.D
        D2 = D * D
        d = (np.einsum("mi,ekjm->ekji", D2, G[:, 0]) + np.einsum("mj,ekmi->ekji", D2, G[:, 2])
             + np.einsum("mk,emji->ekji", D2, G[:, 5]))
        dd = np.diag(D)
        d += 2 * G[:, 1] * dd[None, None, None, :] * dd[None, None, :, None]
        d += 2 * G[:, 4] * dd[None, None, None, :] * dd[None, :, None, None]
        d += 2 * G[:, 3] * dd[None, None, :, None] * dd[None, :, None, None]
        d *= self.lambda0
        if not self.poisson:
            d += self.lambda1 * G[:, 6]  # mass term: lambda1 * GwJ
        diag = np.ascontiguousarray(d.reshape(-1).astype(dtype))
        self.orc.gs_add(self.ogs, diag)
        return (dtype(1) / diag).astype(dtype)


# ------------------------------------------------------------------------------------------ Schwarz
def extended_masked_ids(N, nel_global, brick_lo, brick_n, EToB):
    """masked global ids of the order-(N+2) extended mesh (create_extended_mesh, :643-741): true
    order-(N+2) C0 numbering; element edges/corners and Dirichlet nodes set to 0."""
    from nekrs_b200 import meshgen  # input generator only
    Ne = N + 2
    dummy = meshgen.HexMesh(N=Ne, nel_global=nel_global, brick_lo=brick_lo, brick_n=brick_n, x=None, y=None, z=None,
                            global_ids=None, EToB=None, vertices=None)
    gids = meshgen.global_ids_at_order(dummy, Ne)
    E = int(np.prod(brick_n))
    Npe = (Ne + 1) ** 3
    edge = sem.edge_node_flags(Ne)
    fn = sem.face_nodes(Ne)
    mapB = np.full((E, Npe), 10 ** 9, dtype=np.int64)
    etob = np.asarray(EToB).reshape(E, 6)
    for f in range(6):
        sel = np.flatnonzero(etob[:, f] > 0)
        if sel.size:
            sub = mapB[np.ix_(sel, fn[f])]
            mapB[np.ix_(sel, fn[f])] = np.minimum(sub, etob[sel, f][:, None])
    # gs-min over the extended numbering
    flat = mapB.reshape(-1)
    order = np.argsort(gids, kind="stable")
    sg = gids[order]
    starts = np.flatnonzero(np.r_[True, sg[1:] != sg[:-1]])
    mins = np.minimum.reduceat(flat[order], starts)
    flat[order] = np.repeat(mins, np.diff(np.r_[starts, sg.size]))
    masked = gids.copy().reshape(E, Npe)
    masked[:, edge] = 0
    masked[flat.reshape(E, Npe) == 1] = 0
    return masked.reshape(-1)


class Schwarz:
    """pMGLevel::build / generate_weights / smoothSchwarz."""

    def __init__(self, orc: Orc, lvl: OElliptic, base: OElliptic, ext_ids, options):
        self.orc, self.lvl, self.options = orc, lvl, options
        m = lvl.mesh
        E, N, Nq, Nqe = m.E, m.N, m.Nq, m.Nq + 2
        self.E, self.N, self.Nqe = E, N, Nqe
        self.ogs_ext = sem.Ogs(ext_ids)
        L = self.element_lengths(base)
        Sx = np.zeros((E, Nqe, Nqe))
        Sy, Sz = np.zeros_like(Sx), np.zeros_like(Sx)
        invL = np.zeros((E, Nqe, Nqe, Nqe))
        lookup = [4, 2, 1, 3, 0, 5]
        etob = m.EToB.reshape(E, 6)
        cache = {}
        for e in range(E):
            fbc = [int(etob[e, lookup[i]]) for i in range(6)]
            S, lam = [], []
            for d in range(3):
                key = (fbc[2 * d], fbc[2 * d + 1], L["left"][d][e], L["middle"][d][e], L["right"][d][e])
                if key not in cache:
                    cache[key] = self.matrices_1d(m, *key)
                s, l = cache[key]
                S.append(s)
                lam.append(l)
            # stored row-major [node][mode]
            Sx[e], Sy[e], Sz[e] = S[0], S[1], S[2]
            diag = lam[0][None, None, :] + lam[1][None, :, None] + lam[2][:, None, None]
            with np.errstate(divide="ignore"):
                invL[e] = np.where(diag > 1e-5, 1.0 / diag, 0.0)
        self.Sx, self.Sy, self.Sz = (np.ascontiguousarray(a.reshape(-1).astype(f32)) for a in (Sx, Sy, Sz))
        self.invL = np.ascontiguousarray(invL.reshape(-1).astype(f32))
        self.work1 = np.zeros(E * Nqe ** 3, f32)
        self.work2 = np.zeros(E * Nqe ** 3, f32)
        self.wts = self.generate_weights()

    @staticmethod
    def element_lengths(base: OElliptic):
        m = base.mesh
        E, N, Nq = m.E, m.N, m.Nq
        w = m.gllw
        X = [a.reshape(E, Nq, Nq, Nq) for a in (m.x, m.y, m.z)]  # [e,k,j,i]
        mid = []
        if Nq == 2:
            sl, wt = slice(0, 2), np.ones((2, 2))
        else:
            sl = slice(1, Nq - 1)
            ww = w[0:Nq - 2]  # w[j-1], j = 1..Nq-2
            wt = ww[:, None] * ww[None, :]
        for d in range(3):
            if d == 0:
                dist = [a[:, sl, sl, Nq - 1] - a[:, sl, sl, 0] for a in X]
            elif d == 1:
                dist = [a[:, sl, Nq - 1, sl] - a[:, sl, 0, sl] for a in X]
            else:
                dist = [a[:, Nq - 1, sl, sl] - a[:, 0, sl, sl] for a in X]
            den = dist[0] ** 2 + dist[1] ** 2 + dist[2] ** 2
            l2 = (wt[None] / den).sum(axis=(1, 2)) / wt.sum()
            mid.append(1.0 / np.sqrt(l2))
        if Nq == 2:
            return {"left": mid, "middle": mid, "right": mid}
        l = np.zeros((E, Nq, Nq, Nq))
        s = slice(1, N)
        l[:, s, s, 0] = mid[0][:, None, None]
        l[:, s, s, Nq - 1] = mid[0][:, None, None]
        l[:, s, 0, s] = mid[1][:, None, None]
        l[:, s, Nq - 1, s] = mid[1][:, None, None]
        l[:, 0, s, s] = mid[2][:, None, None]
        l[:, Nq - 1, s, s] = mid[2][:, None, None]
        lf = np.ascontiguousarray(l.reshape(-1))
        base.orc.gs_add(m.ogs, lf)
        l = lf.reshape(E, Nq, Nq, Nq)
        left = [l[:, 1, 1, 0] - mid[0], l[:, 1, 0, 1] - mid[1], l[:, 0, 1, 1] - mid[2]]
        right = [l[:, 1, 1, Nq - 1] - mid[0], l[:, 1, Nq - 1, 1] - mid[1], l[:, Nq - 1, 1, 1] - mid[2]]
        tol = 1e-12
        for d in range(3):
            left[d] = np.where((np.abs(left[d]) < tol) | (left[d] < -tol), mid[d], left[d])
            right[d] = np.where((np.abs(right[d]) < tol) | (right[d] < -tol), mid[d], right[d])
        return {"left": left, "middle": mid, "right": right}

    @staticmethod
    def matrices_1d(m: OMesh, lbc, rbc, ll, lm, lr):
        n = m.N
        nl = n + 3
        D, gw = m.D, m.gllw
        ah = D.T @ (gw[:, None] * D)
        a, b = np.zeros((nl, nl)), np.zeros((nl, nl))
        i0 = 1 if lbc == 1 else 0
        i1 = n - 1 if rbc == 1 else n
        a[1, 1] = 1.0
        a[n + 1, n + 1] = 1.0
        a[i0 + 1:i1 + 2, i0 + 1:i1 + 2] = (2.0 / lm) * ah[i0:i1 + 1, i0:i1 + 1]
        if lbc == 0:
            fac = 2.0 / ll
            a[0, 0] = fac * ah[n - 1, n - 1]
            a[1, 0] = fac * ah[n, n - 1]
            a[0, 1] = fac * ah[n - 1, n]
            a[1, 1] = a[1, 1] + fac * ah[n, n]
        else:
            a[0, 0] = 1.0
        if rbc == 0:
            fac = 2.0 / lr
            a[n + 1, n + 1] = a[n + 1, n + 1] + fac * ah[0, 0]
            a[n + 2, n + 1] = fac * ah[1, 0]
            a[n + 1, n + 2] = fac * ah[0, 1]
            a[n + 2, n + 2] = fac * ah[1, 1]
        else:
            a[n + 2, n + 2] = 1.0
        b[1, 1] = 1.0
        b[n + 1, n + 1] = 1.0
        idx = np.arange(i0, i1 + 1)
        b[idx + 1, idx + 1] = 0.5 * lm * gw[idx]
        if lbc == 0:
            b[0, 0] = 0.5 * ll * gw[n - 1]
            b[1, 1] = b[1, 1] + 0.5 * ll * gw[n]
        else:
            b[0, 0] = 1.0
        if rbc == 0:
            b[n + 1, n + 1] = b[n + 1, n + 1] + 0.5 * lr * gw[0]
            b[n + 2, n + 2] = 0.5 * lr * gw[1]
        else:
            b[n + 2, n + 2] = 1.0
        lam, V = scipy.linalg.eigh(0.5 * (a + a.T), b)  # dsygv itype=1: V^T B V = I, ascending
        S = V.copy()  # S[node][mode]
        if lbc > 0:
            S[0, :] = 0
        if lbc == 1:
            S[1, :] = 0
        if rbc > 0:
            S[nl - 1, :] = 0
        if rbc == 1:
            S[nl - 2, :] = 0
        return S, lam

    def generate_weights(self):
        E, Nq, Nqe = self.E, self.N + 1, self.Nqe
        w1 = np.ones((E, Nqe, Nqe, Nqe), f32)
        w2 = np.ones((E, Nqe, Nqe, Nqe), f32)
        s = slice(1, Nqe - 1)

        def extrude(a1, l1, f1, a2, l2, f2):
            f1, f2 = f32(f1), f32(f2)
            for lo, lo2 in ((l1, l2), (Nqe - l1 - 1, Nqe - l2 - 1)):
                a1[:, s, s, lo] = f1 * a1[:, s, s, lo] + f2 * a2[:, s, s, lo2]
            for lo, lo2 in ((l1, l2), (Nqe - l1 - 1, Nqe - l2 - 1)):
                a1[:, s, lo, s] = f1 * a1[:, s, lo, s] + f2 * a2[:, s, lo2, s]
            for lo, lo2 in ((l1, l2), (Nqe - l1 - 1, Nqe - l2 - 1)):
                a1[:, lo, s, s] = f1 * a1[:, lo, s, s] + f2 * a2[:, lo2, s, s]

        extrude(w2, 0, 0.0, w1, 0, 1.0)
        flat = np.ascontiguousarray(w1.reshape(-1))
        self.orc.gs_add(self.ogs_ext, flat)
        w1 = flat.reshape(E, Nqe, Nqe, Nqe)
        extrude(w1, 0, 1.0, w2, 0, -1.0)
        extrude(w1, 2, 1.0, w1, 0, 1.0)
        wts = np.ascontiguousarray(w1[:, 1:Nq + 1, 1:Nq + 1, 1:Nq + 1].reshape(-1))
        self.orc.gs_add(self.lvl.ogs, wts)
        return (f32(1.0) / wts).astype(f32)

    def smooth(self, u, Su):
        o, E, N = self.orc, self.E, self.N
        o.pre_fdm(E, N, u, self.work1)
        o.gs_add(self.ogs_ext, self.work1)
        if compare(self.options, "MULTIGRID SMOOTHER", "RAS"):
            o.fused_fdm(E, N, Su, self.Sx, self.Sy, self.Sz, self.invL, self.lvl.inv_degree_f, self.work1, 1)
            o.gs_add(self.lvl.ogs, Su)
        else:
            o.fused_fdm(E, N, self.work2, self.Sx, self.Sy, self.Sz, self.invL, self.wts, self.work1, 0)
            o.gs_add(self.ogs_ext, self.work2)
            o.post_fdm(E, N, self.work1, self.work2, Su, self.wts)
            o.gs_add(self.lvl.ogs, Su)
        self.lvl.apply_mask(Su)


# ------------------------------------------------------------------------------------------ MG level
class OLevel:
    def __init__(self, orc, ell: OElliptic, base: OElliptic, degree, is_coarse, options, ext_ids=None,
                 need_smoother=True):
        self.orc, self.ell, self.base, self.degree, self.is_coarse, self.options = orc, ell, base, degree, is_coarse, options
        self.Nrows = ell.mesh.Nlocal
        n = self.Nrows
        self.x, self.rhs, self.res = np.zeros(n, f32), np.zeros(n, f32), np.zeros(n, f32)
        self.s_res, self.s_res2, self.s_upd = np.zeros(n, f32), np.zeros(n, f32), np.zeros(n, f32)
        self.R = None
        self.has_smoother = False
        if need_smoother:
            self.setup_smoother(ext_ids)

    def setup_smoother(self, ext_ids):
        o = self.options
        minM = float(o.get("MULTIGRID CHEBYSHEV MIN EIGENVALUE BOUND FACTOR", 0.1))
        maxM = float(o.get("MULTIGRID CHEBYSHEV MAX EIGENVALUE BOUND FACTOR", 1.1))
        useASM, useRAS = compare(o, "MULTIGRID SMOOTHER", "ASM"), compare(o, "MULTIGRID SMOOTHER", "RAS")
        self.schwarz, self.inv_diag = None, None
        if useASM or useRAS:
            self.smoother_type = "ASM" if useASM else "RAS"
            self.schwarz = Schwarz(self.orc, self.ell, self.base, ext_ids, o)
        else:
            assert compare(o, "MULTIGRID SMOOTHER", "DAMPEDJACOBI"), "Invalid pMGLevel smoother!"
            self.smoother_type = "JACOBI"
            self.inv_diag = self.ell.build_inv_diag(f32)
        self.has_smoother = True
        self.down = self.up = 3
        if compare(o, "MULTIGRID SMOOTHER", "CHEBYSHEV"):
            self.cheby_smoother = self.smoother_type
            self.smoother_type = "CHEBYSHEV"
            rho = self.max_eig()
            self.lambda1, self.lambda0, self.max_eig_value = maxM * rho, minM * rho, rho
            if not self.is_coarse:
                self.down = self.up = int(o.get("MULTIGRID CHEBYSHEV DEGREE", 3))
        if compare(o, "MULTIGRID SMOOTHER", "FOURTHOPT"):
            self.up_betas, self.down_betas = optimal_coeffs(self.up), optimal_coeffs(self.down)
            self.smoother_type = "OPT_FOURTH"
        elif compare(o, "MULTIGRID SMOOTHER", "FOURTH"):
            self.up_betas, self.down_betas = [1.0] * self.up, [1.0] * self.down
            self.smoother_type = "FOURTH"

    # ---- ops
    def Ax(self, x, Ax):
        self.ell.operator(x, Ax)

    def residual(self, rhs, x, res):
        self.ell.operator(x, res)
        self.orc.axpby(self.Nrows, 1.0, rhs, -1.0, res)

    def coarsen(self, x, Rx, fine_inv_degree_f, NfOrder):
        self.orc.axmy(x.size, 1.0, fine_inv_degree_f, x)
        self.orc.coarsen(self.ell.mesh.E, NfOrder, self.degree, self.R, x, Rx)
        self.orc.gs_add(self.ell.ogs, Rx)
        self.ell.apply_mask(Rx)

    def prolongate(self, x, Px, NfOrder):
        self.orc.prolongate(self.ell.mesh.E, NfOrder, self.degree, self.R, x, Px)

    def smoother(self, x, Sx):
        if self.cheby_smoother == "JACOBI":
            self.orc.axmyz(self.Nrows, 1.0, self.inv_diag, x, Sx)
        else:
            self.schwarz.smooth(x, Sx)

    def smooth(self, rhs, x, x_is_zero):
        t = self.smoother_type
        if not x_is_zero and t in ("ASM", "RAS"):
            return
        if t == "CHEBYSHEV":
            self.smooth_chebyshev(rhs, x, x_is_zero)
        elif t in ("OPT_FOURTH", "FOURTH"):
            self.smooth_fourth(rhs, x, x_is_zero)
        elif t in ("ASM", "RAS"):
            self.schwarz.smooth(rhs, x)
        else:
            self.smooth_jacobi(rhs, x, x_is_zero)

    def smooth_jacobi(self, r, x, x_is_zero):
        o, n = self.orc, self.Nrows
        if x_is_zero:
            o.axmyz(n, 1.0, self.inv_diag, r, x)
            return
        res, d = self.s_res, self.s_upd
        self.Ax(x, res)
        o.axpby(n, 1.0, r, -1.0, res)
        o.axmyz(n, 1.0, self.inv_diag, res, d)
        o.axpby(n, 1.0, d, 1.0, x)

    def smooth_chebyshev(self, r, x, x_is_zero):
        o, n = self.orc, self.Nrows
        deg = self.down if x_is_zero else self.up
        if deg == 0:
            return
        theta = f32(0.5 * (self.lambda1 + self.lambda0))
        delta = f32(0.5 * (self.lambda1 - self.lambda0))
        invTheta = f32(1.0 / theta)
        sigma = f32(theta / delta)
        rho_n = f32(1.0 / sigma)
        res, Ad, d = self.s_res, self.s_res2, self.s_upd
        if x_is_zero:
            x[:] = 0
            res[:] = r
        else:
            self.Ax(x, res)
            o.axpby(n, 1.0, r, -1.0, res)
        self.smoother(res, res)
        o.axpby(n, float(invTheta), res, 0.0, d)
        for _ in range(1, deg):
            self.Ax(d, Ad)
            self.smoother(Ad, Ad)
            rhoSave = rho_n
            rho_n = f32(1.0 / (2.0 * float(sigma) - float(rho_n)))
            rCoeff = f32(2.0 * float(rho_n) / float(delta))
            dCoeff = f32(rho_n * rhoSave)
            o.update_chebyshev(n, float(dCoeff), float(rCoeff), Ad, d, res, x)
        o.axpby(n, 1.0, d, 1.0, x)
        self.ell.apply_mask(x)

    def smooth_fourth(self, r, x, x_is_zero):
        o, n = self.orc, self.Nrows
        deg = self.down if x_is_zero else self.up
        betas = self.down_betas if x_is_zero else self.up_betas
        if deg == 0:
            return
        res, Ad, d = self.s_res, self.s_res2, self.s_upd
        rho = f32(self.lambda1)
        if x_is_zero:
            x[:] = 0
            res[:] = r
        else:
            self.Ax(x, res)
            o.axpby(n, 1.0, r, -1.0, res)
        self.smoother(res, Ad)
        coeff = f32(4.0 / (3.0 * float(rho)))
        o.axpby(n, float(coeff), Ad, 0.0, d)
        for k in range(1, deg):
            self.Ax(d, Ad)
            o.update_fourth_chebyshev(n, float(f32(betas[k - 1])), Ad, d, res, x)
            self.smoother(res, Ad)
            dCoeff = f32((2.0 * k - 1.0) / (2.0 * k + 3.0))
            rCoeff = f32((8.0 * k + 4.0) / ((2.0 * k + 3.0) * float(rho)))
            o.axpby(n, float(rCoeff), Ad, float(dCoeff), d)
        o.axpby(n, float(f32(betas[-1])), d, 1.0, x)
        self.ell.apply_mask(x)

    def max_eig(self):
        """Arnoldi(10) on S*A (ellipticMultiGridLevelSetup.cpp:292-453), deterministic start vector."""
        o, ell, m = self.orc, self.ell, self.ell.mesh
        M = self.Nrows
        k = int(min(10, m.E * m.Np))
        H = np.zeros((k, k))
        Vx = id_uniform(m.global_ids)
        o.gs_add(m.ogs, Vx)
        Vx[ell.mask_ids] = 0.0
        w = ell.inv_degree
        V = [np.zeros(M) for _ in range(k + 1)]
        norm_vo = np.sqrt(o.weighted_inner_prod(M, w, Vx, Vx))
        o.axpby(M, 1.0 / norm_vo, Vx, 0.0, V[0])
        vf, avf = np.zeros(M, f32), np.zeros(M, f32)
        for j in range(k):
            o.copy_d2f(V[j], vf)
            ell.operator(vf, avf)
            self.smoother(avf, vf)
            o.copy_f2d(vf, V[j + 1])
            for i in range(j + 1):
                hij = o.weighted_inner_prod(M, w, V[i], V[j + 1])
                o.axpby(M, -hij, V[i], 1.0, V[j + 1])
                H[i, j] = hij
            if j + 1 < k:
                nv = np.sqrt(o.weighted_inner_prod(M, w, V[j + 1], V[j + 1]))
                V[j + 1] *= 1.0 / nv
                H[j + 1, j] = nv
        return float(np.max(np.abs(np.linalg.eigvals(H))))


class CoarseJPCG:
    """Coarse-solve stand-in shared with the product (coarse.cu coarseSolver_t): Jacobi-PCG in the
    Chronopoulos-Gear form on the ASSEMBLED N=1 operator (the matrix of ellipticBuildFEMHex3D,
    MG/ellipticBuildFEM.cpp:73-305), unknowns = unique unmasked nodes.

    The element matrices are obtained here by applying the oracle's matrix-free N=1 operator to the
    eight unit vectors (an independent route to the same matrix the product assembles from the
    closed-form entries)."""

    def __init__(self, orc, lvl: OLevel, max_iter, tol):
        import scipy.sparse as sp
        self.orc, self.lvl, self.max_iter, self.tol = orc, lvl, max_iter, tol
        ell, m = lvl.ell, lvl.ell.mesh
        E, Np = m.E, m.Np
        ogs = ell.ogs
        NT = ogs.Ngather
        t_index = np.full(m.Nlocal, -1, dtype=np.int64)
        counts = np.diff(ogs.offsets)
        t_index[ogs.gather_ids] = np.repeat(np.arange(NT), counts)
        self.row_node = ogs.gather_ids[ogs.offsets[:-1]]
        self.t_index = t_index
        Ae = np.zeros((E, Np, Np))
        for mm in range(Np):
            q = np.zeros((E, Np))
            q[:, mm] = 1.0
            out = np.zeros(E * Np)
            orc.ax(m.N, m.element_list, m.ggeo, m.D, np.ascontiguousarray(q.reshape(-1)), out)
            Ae[:, :, mm] = out.reshape(E, Np)
        tn = t_index.reshape(E, Np)
        rows = np.repeat(tn[:, :, None], Np, axis=2).reshape(-1)
        cols = np.repeat(tn[:, None, :], Np, axis=1).reshape(-1)
        keep = (rows >= 0) & (cols >= 0)
        A = sp.coo_matrix((Ae.reshape(-1)[keep], (rows[keep], cols[keep])), shape=(NT, NT)).tocsr()
        A.sum_duplicates()
        self.A = A.astype(f32)
        self.inv_diag = (f32(1.0) / self.A.diagonal()).astype(f32)
        self.NT = NT
        self.last_iter = 0

    def spmv_dots(self, r, u):
        w = (self.A @ u).astype(f32)
        ud = u.astype(np.float64)
        return w, float(np.sum(r.astype(np.float64) * ud)), float(np.sum(w.astype(np.float64) * ud))

    def solve(self, rhs, xE):
        b = rhs[self.row_node].astype(f32)
        x = np.zeros(self.NT, f32)
        r = b.copy()
        u = (self.inv_diag * r).astype(f32)
        p, s = np.zeros(self.NT, f32), np.zeros(self.NT, f32)
        w, gamma, delta = self.spmv_dots(r, u)
        gamma0 = gamma
        beta = 0.0
        alpha = gamma / delta if delta > 0.0 else 0.0
        it = 0
        for it in range(1, self.max_iter + 1):
            a32, b32 = f32(alpha), f32(beta)
            p = (u + b32 * p).astype(f32)
            s = (w + b32 * s).astype(f32)
            x = (x + a32 * p).astype(f32)
            r = (r - a32 * s).astype(f32)
            u = (self.inv_diag * r).astype(f32)
            w, gn, delta = self.spmv_dots(r, u)
            beta = gn / gamma if gamma > 0.0 else 0.0
            den = delta - beta * gn / alpha if alpha != 0.0 else delta
            alpha = gn / den if den > 0.0 else 0.0
            gamma = gn
            if it % 8 == 0 or it == self.max_iter:
                if not (gamma > self.tol * self.tol * gamma0):
                    break
        self.last_iter = min(it, self.max_iter)
        xE[:] = 0
        sel = self.t_index >= 0
        xE[sel] = x[self.t_index[sel]]


# ------------------------------------------------------------------------------------------ solver
class OSolver:
    """ellipticSolveSetup + ellipticSolve for the pressure solve on one rank."""

    def __init__(self, hexmesh, options: dict, orc: Orc = None, poisson=True, lambda0=1.0, lambda1=0.0):
        from nekrs_b200 import meshgen  # mesh/input generator (numbering at the level orders)
        self.orc = orc or Orc()
        self.options = {k.upper(): str(v).upper() for k, v in options.items()}
        self.hex = hexmesh
        o = self.options
        m = OMesh(self.orc, hexmesh.N, hexmesh.Nelements, hexmesh.x, hexmesh.y, hexmesh.z, hexmesh.global_ids,
                  hexmesh.EToB)
        self.mesh = m
        self.nvec = m.Nlocal
        self.ell = OElliptic(self.orc, m, o, poisson=poisson, lambda0=lambda0, lambda1=lambda1)
        if not poisson or lambda0 != 1.0:
            assert not compare(o, "PRECONDITIONER", "MULTIGRID"), "the multigrid restatement is Poisson-only"
        per = 1024 // 8
        self.fieldOffset = ((m.Nlocal + per - 1) // per) * per
        self.levels = []
        self.res_history = []
        self.Niter = 0
        self.proj = None
        if compare(o, "PRECONDITIONER", "MULTIGRID"):
            self.setup_mg(meshgen)
        elif compare(o, "PRECONDITIONER", "JACOBI"):
            self.inv_diag = self.ell.build_inv_diag(np.float64)
        if compare(o, "INITIAL GUESS", "PROJECTION"):
            self.proj = Projection(self, compare(o, "INITIAL GUESS", "PROJECTION-ACONJ"),
                                   int(o.get("RESIDUAL PROJECTION VECTORS", 8)),
                                   int(o.get("RESIDUAL PROJECTION START", 5)))

    def setup_mg(self, meshgen):
        from nekrs_b200.elliptic import mg_level_orders  # pure table lookup of determineMGLevels
        o, hx = self.options, self.hex
        orders = mg_level_orders(o, hx.N)
        coarse_solve = compare(o, "MULTIGRID COARSE SOLVE", "TRUE")
        and_smooth = compare(o, "MULTIGRID COARSE SOLVE AND SMOOTH", "TRUE")
        schwarz = compare(o, "MULTIGRID SMOOTHER", "ASM") or compare(o, "MULTIGRID SMOOTHER", "RAS")
        for n, Nc in enumerate(orders):
            is_coarse = n == len(orders) - 1
            if Nc == hx.N:
                mesh = self.mesh
            else:
                xc = sem.interpolate_nodes(hx.x, hx.N, Nc)
                yc = sem.interpolate_nodes(hx.y, hx.N, Nc)
                zc = sem.interpolate_nodes(hx.z, hx.N, Nc)
                mesh = OMesh(self.orc, Nc, hx.Nelements, xc, yc, zc, meshgen.global_ids_at_order(hx, Nc), hx.EToB)
            ell = OElliptic(self.orc, mesh, o)
            need = (not is_coarse) or len(orders) == 1 or (not coarse_solve) or and_smooth
            ext = extended_masked_ids(Nc, hx.nel_global, hx.brick_lo, hx.brick_n, hx.EToB) if (schwarz and need) else None
            lvl = OLevel(self.orc, ell, self.ell, Nc, is_coarse, o, ext, need)
            if n > 0:
                Nf = orders[n - 1]
                gf, _ = sem.jacobi_gll(Nf)
                gc, _ = sem.jacobi_gll(Nc)
                lvl.R = np.ascontiguousarray(sem.interpolation_matrix_1d(gc, gf).T.astype(f32))
                lvl.Nf = Nf
            self.levels.append(lvl)
        self.coarse = None
        if coarse_solve:
            self.coarse = CoarseJPCG(self.orc, self.levels[-1], int(o.get("COARSE SOLVER MAXIMUM ITERATIONS", 200)),
                                     float(o.get("COARSE SOLVER TOLERANCE", 1e-1)))
        self.and_smooth = and_smooth

    # ---- V-cycle (MGSolver.cpp:167-193)
    def vcycle(self, k):
        lv = self.levels[k]
        if k == len(self.levels) - 1:
            self.coarse_solve(lv.rhs, lv.x)
            return
        lc = self.levels[k + 1]
        lv.smooth(lv.rhs, lv.x, True)
        lv.residual(lv.rhs, lv.x, lv.res)
        lc.coarsen(lv.res, lc.rhs, lv.ell.inv_degree_f, lv.degree)
        self.vcycle(k + 1)
        lc.prolongate(lc.x, lv.x, lv.degree)
        lv.smooth(lv.rhs, lv.x, False)

    # ---- additive cycle (MGSolver.cpp:195-251 with coarsenV / schwarzSolve / prolongateV, :32-78): restrict the
    #      right-hand side to every level, smooth every level from zero, solve the coarse problem, add the prolongated
    #      corrections.  (The reference runs the coarse solve in a second OpenMP task: same arithmetic.)
    def additive_vcycle(self):
        L = self.levels
        for k in range(len(L) - 1):  # coarsenV
            L[k].res[:] = L[k].rhs
            L[k + 1].coarsen(L[k].res, L[k + 1].rhs, L[k].ell.inv_degree_f, L[k].degree)
        for k in range(len(L) - 1):  # schwarzSolve
            L[k].smooth(L[k].rhs, L[k].x, True)
            L[k].res[:] = L[k].rhs
            L[k + 1].coarsen(L[k].res, L[k + 1].rhs, L[k].ell.inv_degree_f, L[k].degree)
        self.coarse_solve(L[-1].rhs, L[-1].x)
        for k in range(len(L) - 2, -1, -1):  # prolongateV
            L[k + 1].prolongate(L[k + 1].x, L[k].x, L[k].degree)

    def coarse_solve(self, rhs, x):
        base = self.levels[-1]
        if self.coarse is None:
            base.smooth(rhs, x, True)
        elif self.and_smooth:
            base.smooth(rhs, x, True)
            base.residual(rhs, x, base.res)
            tmp = base.s_upd
            self.coarse.solve(base.res, tmp)
            self.orc.axpby(base.Nrows, 1.0, tmp, 1.0, x)
            base.smooth(rhs, x, False)
        else:
            self.coarse.solve(rhs, x)

    def preconditioner(self, r, z):
        o, orc, n = self.options, self.orc, self.nvec
        if compare(o, "PRECONDITIONER", "JACOBI"):
            orc.axmyz(n, 1.0, r, self.inv_diag, z)
        elif compare(o, "PRECONDITIONER", "MULTIGRID"):
            l0 = self.levels[0]
            l0.x[:] = 0
            orc.copy_d2f(np.ascontiguousarray(r[:n]), l0.rhs)
            if compare(o, "MGSOLVER CYCLE", "ADDITIVE"):
                self.additive_vcycle()
            else:
                self.vcycle(0)
            zz = np.zeros(n)
            orc.copy_f2d(l0.x, zz)
            z[:n] = zz
        else:
            z[:] = r
        if self.ell.allNeumann:
            self.zero_mean(z)

    def zero_mean(self, q):
        n = self.mesh.Nlocal
        mean = self.orc.sum(n, q) / float(n)
        q[:n] += -mean

    def wnorm(self, v):
        n = self.nvec
        return np.sqrt(self.orc.weighted_norm2_sq(n, self.ell.inv_degree, v)) * np.sqrt(self.resNormFactor)

    # ---- ellipticSolve
    def solve(self, rhs, x):
        o, orc, ell = self.options, self.orc, self.ell
        n = self.nvec
        r = np.ascontiguousarray(rhs, dtype=np.float64).copy()
        x = np.ascontiguousarray(x, dtype=np.float64)
        maxIter = int(o.get("MAXIMUM ITERATIONS", 999))
        self.resNormFactor = 1.0 / self.mesh.volume
        self.res_history = []
        Ap = np.zeros(n)
        ell.ax(x, Ap)
        orc.axpby(n, -1.0, Ap, 1.0, r)
        if ell.allNeumann:
            self.zero_mean(r)
        ell.apply_mask(r)
        ell.gs(r)
        x0 = x.copy()
        x[:] = 0
        if self.proj is not None:
            self.res00Norm = self.wnorm(r)
            self.proj.pre(r)
        self.res0Norm = self.wnorm(r)
        tol = float(o.get("SOLVER TOLERANCE", 1e-6))
        if compare(o, "LINEAR SOLVER STOPPING CRITERION", "RELATIVE"):
            tol *= self.res0Norm
        self.resNorm = self.res0Norm
        if compare(o, "SOLVER", "PCG"):
            self.Niter = self.pcg(r, x, tol, maxIter)
        else:
            self.Niter = self.pgmres(r, x, tol, maxIter)
        if self.proj is not None:
            self.proj.post(x)
        else:
            self.res00Norm = self.res0Norm
        orc.axpby(n, 1.0, x0, 1.0, x)
        if ell.allNeumann:
            self.zero_mean(x)
        return x

    def pcg(self, r, x, tol, MAXIT):
        o, orc, ell = self.options, self.orc, self.ell
        n = self.nvec
        flexible = compare(o, "SOLVER", "FLEXIBLE")
        precond = not compare(o, "PRECONDITIONER", "NONE")
        p, Ap = np.zeros(n), np.zeros(n)
        z = np.zeros(n) if precond else r
        w = ell.inv_degree
        rdotr = self.resNorm
        rdotz1, alpha = 0.0, 0.0
        it = 0
        while True:
            it += 1
            rdotz2 = rdotz1
            if precond:
                self.preconditioner(r, z)
                rdotz1 = orc.weighted_inner_prod(n, w, r, z)
            else:
                rdotz1 = rdotr  # PCG.cpp:131 (the norm, literally)
            beta = 0.0
            if it > 1:
                beta = rdotz1 / rdotz2
                if flexible:
                    zdotAp = orc.weighted_inner_prod(n, w, z, Ap)
                    beta = -alpha * zdotAp / rdotz2
            orc.axpby(n, 1.0, z, beta, p)
            ell.operator(p, Ap)
            pAp = orc.weighted_inner_prod(n, w, p, Ap)
            alpha = rdotz1 / (pAp + 1e-300)
            rr = orc.update_pcg(n, w, Ap, alpha, r)
            orc.axpby(n, alpha, p, 1.0, x)
            rdotr = np.sqrt(rr * self.resNormFactor)
            self.res_history.append(rdotr)
            if not (rdotr > tol and it < MAXIT):
                break
        self.resNorm = rdotr
        return it

    def pgmres(self, r, x, tol, MAXIT):
        o, orc, ell = self.options, self.orc, self.ell
        n, fo = self.mesh.Nlocal, self.mesh.Nlocal
        m = int(o.get("PGMRES RESTART", 15))
        flexible = compare(o, "SOLVER", "FLEXIBLE")
        wgt = ell.inv_degree
        V = np.zeros(m * fo)
        Z = np.zeros((m if flexible else 1) * fo)
        w, Ax = np.zeros(n), np.zeros(n)
        b = r.copy()
        H = np.zeros((m + 1, m + 1))  # H[k, i]
        sn, cs, s, y = np.zeros(m), np.zeros(m), np.zeros(m + 1), np.zeros(m)
        rnf = np.sqrt(self.resNormFactor)
        nr = self.resNorm / rnf
        error = self.resNorm

        def update(size):
            for k in range(size - 1, -1, -1):
                y[k] = s[k]
                for mm in range(k + 1, size):
                    y[k] -= H[k, mm] * y[mm]
                y[k] /= H[k, k]
            if flexible:
                orc.update_pgmres_solution(n, fo, size, y, Z, x)
            else:
                zt = np.zeros(n)
                orc.update_pgmres_solution(n, fo, size, y, V, zt)
                tmp = np.zeros(n)
                self.preconditioner(zt, tmp)
                orc.axpby(n, 1.0, tmp, 1.0, x)

        it = 0
        while it < MAXIT:
            s[0] = nr
            v0 = V[0:n]
            orc.axpby(n, 1.0 / nr, r, 0.0, v0)
            done = False
            for i in range(m):
                Mv = Z[i * fo:i * fo + n] if flexible else Z[0:n]
                vi = np.ascontiguousarray(V[i * fo:i * fo + n])
                self.preconditioner(vi, Mv)
                ell.operator(np.ascontiguousarray(Mv), w)
                yy = orc.weighted_inner_prod_multi(n, i + 1, fo, wgt, V, w)
                y[:i + 1] = yy
                nw = np.sqrt(orc.gram_schmidt(n, fo, i + 1, wgt, y, V, w))
                H[i + 1, i] = nw
                if i < m - 1:
                    orc.axpby(n, 1.0 / nw, w, 0.0, V[(i + 1) * fo:(i + 1) * fo + n])
                H[:i + 1, i] = y[:i + 1]
                for k in range(i):
                    h1, h2 = H[k, i], H[k + 1, i]
                    H[k, i] = cs[k] * h1 + sn[k] * h2
                    H[k + 1, i] = -sn[k] * h1 + cs[k] * h2
                h1, h2 = H[i, i], H[i + 1, i]
                hr = np.sqrt(h1 * h1 + h2 * h2)
                cs[i], sn[i] = h1 / hr, h2 / hr
                H[i, i] = cs[i] * h1 + sn[i] * h2
                H[i + 1, i] = 0
                s[i + 1] = -sn[i] * s[i]
                s[i] = cs[i] * s[i]
                it += 1
                error = abs(s[i + 1]) * rnf
                self.res_history.append(error)
                if error < tol or it == MAXIT:
                    update(i + 1)
                    done = True
                    break
            if done:
                break
            update(m)
            ell.operator(x, Ax)
            nr = np.sqrt(orc.fused_residual_and_norm(n, wgt, b, Ax, r))
            error = nr * rnf
            if error <= tol:
                break
        self.resNorm = error
        return it


class Projection:
    """SolutionProjection (ellipticSolutionProjection.cpp:44-288), Nfields = 1."""

    def __init__(self, solver: OSolver, aconj, max_vecs, n_steps):
        self.s, self.aconj, self.max_vecs, self.n_steps = solver, aconj, max_vecs, n_steps
        n = solver.mesh.Nlocal
        self.n = n
        self.num, self.timestep = 0, 0
        self.xx = np.zeros(max_vecs * n)
        self.bb = np.zeros((1 if aconj else max_vecs) * n)
        self.xbar = np.zeros(n)
        self.alpha = np.zeros(max_vecs)
        self.w = solver.mesh.ogs.inv_degree

    def matvec(self, Ax_off, x_off):
        n = self.n
        xin = np.ascontiguousarray(self.xx[x_off * n:(x_off + 1) * n])
        out = np.zeros(n)
        self.s.ell.operator(xin, out)
        self.bb[Ax_off * n:(Ax_off + 1) * n] = out

    def update_space(self):
        if self.num <= 0:
            return
        n, m, orc = self.n, self.num, self.s.orc
        yv = np.ascontiguousarray(self.bb[(0 if self.aconj else (m - 1) * n):][:n])
        self.alpha[:m] = orc.weighted_inner_prod_multi(n, m, n, self.w, self.xx, yv)
        norm_orig = self.alpha[m - 1]
        for arr in ([self.xx] if self.aconj else [self.xx, self.bb]):
            dst = arr[(m - 1) * n:m * n]
            for k in range(m - 1):
                dst[:] = -self.alpha[k] * arr[k * n:(k + 1) * n] + 1.0 * dst
        norm_new = np.sqrt(norm_orig - np.sum(self.alpha[:m - 1] ** 2))
        if norm_new / norm_orig > 1e-7:
            sc = 1.0 / norm_new
            self.xx[(m - 1) * n:m * n] *= sc
            if not self.aconj:
                self.bb[(m - 1) * n:m * n] *= sc
        else:
            self.num -= 1

    def pre(self, r):
        self.timestep += 1
        if self.timestep < self.n_steps or self.num <= 0:
            return
        n, m, orc = self.n, self.num, self.s.orc
        self.alpha[:m] = orc.weighted_inner_prod_multi(n, m, n, self.w, self.xx, np.ascontiguousarray(r[:n]))
        self.xbar[:] = self.alpha[0] * self.xx[:n]
        for k in range(1, m):
            self.xbar += self.alpha[k] * self.xx[k * n:(k + 1) * n]
        if not self.aconj:
            rt = self.alpha[0] * self.bb[:n]
            for k in range(1, m):
                rt += self.alpha[k] * self.bb[k * n:(k + 1) * n]
            orc.axpby(n, -1.0, rt, 1.0, r)
        else:
            out = np.zeros(n)
            self.s.ell.operator(self.xbar, out)
            self.bb[:n] = out
            orc.axpby(n, -1.0, out, 1.0, r)

    def post(self, x):
        if self.timestep < self.n_steps:
            return
        n, orc = self.n, self.s.orc
        if self.num == 0:
            self.num = 1
            self.xx[:n] = x[:n]
        elif self.num == self.max_vecs:
            self.num = 1
            orc.axpby(n, 1.0, self.xbar, 1.0, x)
            self.xx[:n] = x[:n]
        else:
            self.num += 1
            self.xx[(self.num - 1) * n:self.num * n] = x[:n]
            orc.axpby(n, 1.0, self.xbar, 1.0, x)
        prev = self.num
        self.matvec(0 if self.aconj else self.num - 1, self.num - 1)
        self.update_space()
        if self.num < prev:
            self.num = 1
            self.xx[:n] = x[:n]
            self.matvec(0, 0)
            self.update_space()

# ------------------------------------------------------------------------------------------ block solver
class OBlockElliptic:
    """elliptic_t with Nfields = 3 (ellipticSetup.cpp:81-131,192-249; ellipticOgs.cpp:17-131): boundary flags and mask
    per field, the UNMASKED mesh numbering for the gather-scatter of all fields, block Helmholtz or stress operator."""

    def __init__(self, orc: Orc, mesh: OMesh, EToB, lambda0, lambda1, offset, stress_form):
        self.orc, self.mesh, self.offset, self.stress = orc, mesh, offset, bool(stress_form)
        self.Nfields = 3
        etob = np.asarray(EToB, dtype=np.int32).reshape(self.Nfields, mesh.E * 6)
        ids = []
        for f in range(self.Nfields):
            mi, _ = sem.dirichlet_mask_ids(mesh.N, mesh.E, etob[f], mesh.ogs, orc)
            ids.append(np.asarray(mi, dtype=np.int64) + f * offset)
        self.mask_ids = np.concatenate(ids).astype(np.int32)
        self.ogs = mesh.ogs
        self.allNeumann = 0
        w = np.zeros(self.Nfields * offset)
        for f in range(self.Nfields):
            w[f * offset:f * offset + mesh.Nlocal] = mesh.ogs.inv_degree
        self.inv_degree = w
        # one constant per field, read at lambda[fld * loffset] with loffset = 1
        self.lam0 = np.ascontiguousarray(lambda0, dtype=np.float64)
        self.lam1 = np.ascontiguousarray(lambda1, dtype=np.float64)
        if self.stress:
            self.vgeo = volume_factors(mesh)

    def ax(self, q, Aq):
        m = self.mesh
        if self.stress:
            self.orc.ax_stress(m.N, m.element_list, self.vgeo, m.D, q, Aq, self.lam0, self.lam1, self.offset, 1)
        else:
            self.orc.ax_block(m.N, m.element_list, m.ggeo, m.D, q, Aq, self.lam0, self.lam1, self.offset, 1)

    def apply_mask(self, v):
        self.orc.mask(self.mask_ids, v)

    def gs(self, v):
        self.orc.gs_add(self.ogs, v, k=self.Nfields, stride=self.offset)

    def operator(self, q, Aq, masked=True):
        self.ax(q, Aq)
        if masked:
            self.apply_mask(Aq)
        self.gs(Aq)

    def build_inv_diag(self):
        """ellipticBlockBuildDiagonalHex3D per field (also in stress form: ellipticUpdateJacobi.cpp:38-47), gs, 1/x."""
        from . import kernels as K
        m = self.mesh
        out = np.zeros(self.Nfields * self.offset)
        g = m.ggeo.reshape(m.E, 7, m.Np)
        for f in range(self.Nfields):
            d = K.build_diagonal(m.N, m.E, g, m.D, np.array([self.lam0[f]]), np.array([self.lam1[f]]), poisson=False,
                                 lambda_field=False)
            out[f * self.offset:f * self.offset + m.Nlocal] = d
        self.gs(out)
        inv = np.zeros_like(out)
        for f in range(self.Nfields):
            sl = slice(f * self.offset, f * self.offset + m.Nlocal)
            inv[sl] = 1.0 / out[sl]
        return inv


def volume_factors(m: OMesh):
    """mesh->vgeo (meshGeometricFactorsHex3D.cpp; ids mesh3D.h:82-93): [E][12][Np] = rx,ry,rz,sx,sy,sz,tx,ty,tz,J,JW,1/JW."""
    E, Nq, Np, D = m.E, m.Nq, m.Np, m.D
    X = [a.reshape(E, Nq, Nq, Nq) for a in (m.x, m.y, m.z)]
    J = np.empty((E, Nq, Nq, Nq, 3, 3))
    for a, f in enumerate(X):
        J[..., a, 0] = np.einsum("im,ekjm->ekji", D, f)
        J[..., a, 1] = np.einsum("jm,ekmi->ekji", D, f)
        J[..., a, 2] = np.einsum("km,emji->ekji", D, f)
    Ji = np.linalg.inv(J)  # rows r,s,t ; columns x,y,z
    det = np.linalg.det(J)
    w = m.gllw
    W = (w[:, None, None] * w[None, :, None] * w[None, None, :]).reshape(1, Np)
    v = np.zeros((E, 12, Np))
    for a in range(3):
        for b in range(3):
            v[:, 3 * a + b] = Ji[..., a, b].reshape(E, Np)
    v[:, 9] = det.reshape(E, Np)
    v[:, 10] = det.reshape(E, Np) * W
    v[:, 11] = 1.0 / v[:, 10]
    return v


class OBlockSolver(OSolver):
    """ellipticSolveSetup + ellipticSolve for a block (velocity-type) solve: Jacobi or no preconditioner, PCG."""

    def __init__(self, hexmesh, options: dict, EToB, lambda0, lambda1, orc: Orc = None, stress_form=False):
        self.orc = orc or Orc()
        self.options = {k.upper(): str(v).upper() for k, v in options.items()}
        self.hex = hexmesh
        o = self.options
        assert compare(o, "SOLVER", "PCG") and not compare(o, "PRECONDITIONER", "MULTIGRID")
        m = OMesh(self.orc, hexmesh.N, hexmesh.Nelements, hexmesh.x, hexmesh.y, hexmesh.z, hexmesh.global_ids,
                  np.asarray(EToB).reshape(3, -1)[0])
        self.mesh = m
        per = 1024 // 8
        self.fieldOffset = ((m.Nlocal + per - 1) // per) * per
        self.nvec = 3 * self.fieldOffset
        self.ell = OBlockElliptic(self.orc, m, EToB, lambda0, lambda1, self.fieldOffset, stress_form)
        self.levels, self.res_history, self.Niter, self.proj = [], [], 0, None
        if compare(o, "PRECONDITIONER", "JACOBI"):
            self.inv_diag = self.ell.build_inv_diag()

"""numpy restatement of the reference's host-side *setup* for the Poisson path.

TEST INFRASTRUCTURE ONLY (see oracle/nrs_oracle.c header).

Covers (reference file:line in each docstring): GLL nodes/weights, derivative
and interpolation matrices, geometric factors, ogs gather maps + invDegree,
Dirichlet masks, element lists, pMG level meshes.
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial import legendre as npleg


# --------------------------------------------------------------------------- basis
def jacobi_gll(N: int):
    """GLL nodes and weights (src/mesh/meshBasis1D.cpp:237-266 `JacobiGLL`).

    The reference gets the interior nodes as eigenvalues of the Jacobi(1,1)
    matrix and the weights by mass lumping; both equal the classical GLL rule
    to round-off.  Restated with the closed forms: roots of P_N' and
    w_i = 2 / (N (N+1) P_N(x_i)^2).
    """
    if N == 1:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    c = np.zeros(N + 1)
    c[N] = 1.0
    dc = npleg.legder(c)
    xi = np.sort(npleg.legroots(dc))
    # polish with Newton on P_N'
    d2c = npleg.legder(dc)
    for _ in range(3):
        xi = xi - npleg.legval(xi, dc) / npleg.legval(xi, d2c)
    x = np.concatenate([[-1.0], xi, [1.0]])
    x = 0.5 * (x - x[::-1])
    w = 2.0 / (N * (N + 1) * npleg.legval(x, c) ** 2)
    return x, w


def dmatrix_1d(x: np.ndarray) -> np.ndarray:
    """D[i][m] = l_m'(x_i)  (meshBasis1D.cpp:99-119 `Dmatrix1D`, Vr/V).  Row-major.

    Restated through barycentric weights (same matrix, better conditioned than
    the Vandermonde solve the reference uses)."""
    n = len(x)
    dx = x[:, None] - x[None, :]
    np.fill_diagonal(dx, 1.0)
    bw = 1.0 / np.prod(dx, axis=1)
    D = (bw[None, :] / bw[:, None]) / dx
    np.fill_diagonal(D, 0.0)
    np.fill_diagonal(D, -np.sum(D, axis=1))
    return D


def interpolation_matrix_1d(x_in: np.ndarray, x_out: np.ndarray) -> np.ndarray:
    """I[o][i] = l_i(x_out[o])  (meshBasis1D.cpp:131-150 `InterpolationMatrix1D`)."""
    n = len(x_in)
    dx = x_in[:, None] - x_in[None, :]
    np.fill_diagonal(dx, 1.0)
    bw = 1.0 / np.prod(dx, axis=1)
    I = np.zeros((len(x_out), n))
    for o, xo in enumerate(x_out):
        d = xo - x_in
        hit = np.flatnonzero(np.abs(d) < 1e-14)
        if hit.size:
            I[o, hit[0]] = 1.0
        else:
            t = bw / d
            I[o] = t / t.sum()
    return I


def face_nodes(N: int) -> np.ndarray:
    """faceNodes[6][Nfp]  (meshBasisHex3D.cpp:53-82): f0 t=-1, f1 s=-1, f2 r=+1,
    f3 s=+1, f4 r=-1, f5 t=+1; nodes in increasing node-index order."""
    Nq = N + 1
    n = np.arange(Nq ** 3)
    i, j, k = n % Nq, (n // Nq) % Nq, n // (Nq * Nq)
    return np.stack([n[k == 0], n[j == 0], n[i == N], n[j == N], n[i == 0], n[k == N]])


def edge_node_flags(N: int) -> np.ndarray:
    """bool[Np]: node lies on one of the 12 element edges
    (meshLoadReferenceNodesHex3D.cpp:126-150 `edgeNodes`)."""
    Nq = N + 1
    n = np.arange(Nq ** 3)
    i, j, k = n % Nq, (n // Nq) % Nq, n // (Nq * Nq)
    bi, bj, bk = (i == 0) | (i == N), (j == 0) | (j == N), (k == 0) | (k == N)
    return (bi & bj) | (bi & bk) | (bj & bk)


# --------------------------------------------------------------------------- geometry
def geometric_factors(x, y, z, N):
    """ggeo[E][7][Np] (G00,G01,G11,G12,G02,G22,GWJ) and J[E][Np]
    (kernels/mesh/geometricFactorsHex3D.okl:26-142)."""
    Nq = N + 1
    g, w = jacobi_gll(N)
    D = dmatrix_1d(g)
    X = [np.asarray(a, dtype=np.float64).reshape(-1, Nq, Nq, Nq) for a in (x, y, z)]  # [e,k,j,i]
    dr = [np.einsum("im,ekjm->ekji", D, a) for a in X]
    ds = [np.einsum("jm,ekmi->ekji", D, a) for a in X]
    dt = [np.einsum("km,emji->ekji", D, a) for a in X]
    xr, yr, zr = dr
    xs, ys, zs = ds
    xt, yt, zt = dt
    J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt)
    Ji = 1.0 / J
    JW = J * w[None, None, None, :] * w[None, None, :, None] * w[None, :, None, None]
    rx, ry, rz = (ys * zt - zs * yt) * Ji, -(xs * zt - zs * xt) * Ji, (xs * yt - ys * xt) * Ji
    sx, sy, sz = -(yr * zt - zr * yt) * Ji, (xr * zt - zr * xt) * Ji, -(xr * yt - yr * xt) * Ji
    tx, ty, tz = (yr * zs - zr * ys) * Ji, -(xr * zs - zr * xs) * Ji, (xr * ys - yr * xs) * Ji
    E = J.shape[0]
    gg = np.empty((E, 7, Nq, Nq, Nq))
    gg[:, 0] = JW * (rx * rx + ry * ry + rz * rz)
    gg[:, 1] = JW * (rx * sx + ry * sy + rz * sz)
    gg[:, 4] = JW * (rx * tx + ry * ty + rz * tz)
    gg[:, 2] = JW * (sx * sx + sy * sy + sz * sz)
    gg[:, 3] = JW * (sx * tx + sy * ty + sz * tz)
    gg[:, 5] = JW * (tx * tx + ty * ty + tz * tz)
    gg[:, 6] = JW
    return gg.reshape(E, 7, Nq ** 3), J.reshape(E, Nq ** 3)


def interpolate_nodes(xf, Nf: int, Nc: int):
    """Coarse-level node coordinates: tensor interpolation of the order-Nf nodes to
    the order-Nc GLL points (meshPhysicalNodesHex3D.cpp:46-50 `map_m_to_n`,
    called from createMeshMG, meshSetup.cpp:338)."""
    gf, _ = jacobi_gll(Nf)
    gc, _ = jacobi_gll(Nc)
    I = interpolation_matrix_1d(gf, gc)  # [Nqc, Nqf]
    a = np.asarray(xf).reshape(-1, Nf + 1, Nf + 1, Nf + 1)
    a = np.einsum("ai,ekji->ekja", I, a)
    a = np.einsum("bj,ekja->ekba", I, a)
    a = np.einsum("ck,ekba->ecba", I, a)
    return a.reshape(-1)


# --------------------------------------------------------------------------- gather-scatter
class Ogs:
    """Single-rank restatement of `ogsSetup` (3rd_party/gslib/ogs/src/ogsSetup.cpp:111-397).

    ids == 0 are ignored (ogs.hpp:42-44).  Rows: one per distinct id, ordered by
    the smallest local index of the row (ogsSetup.cpp:222-233); inside a row the
    local indices ascend (sort by baseId then localId, :196; re-sorted by
    localId, :212).  Singleton rows are kept in the CSR, as in the reference.
    invDegree[n] = 1/(row length), 1 for ignored nodes (:366-392).
    """

    def __init__(self, ids: np.ndarray):
        ids = np.asarray(ids, dtype=np.int64)
        self.N = ids.size
        nz = np.flatnonzero(ids != 0)
        order = nz[np.argsort(ids[nz], kind="stable")]
        sid = ids[order]
        if order.size:
            starts = np.flatnonzero(np.r_[True, sid[1:] != sid[:-1]])
        else:
            starts = np.zeros(0, dtype=np.int64)
        counts = np.diff(np.r_[starts, order.size])
        first_local = order[starts] if order.size else starts
        perm = np.argsort(first_local, kind="stable")
        cnt_p = counts[perm]
        offsets = np.zeros(perm.size + 1, dtype=np.int64)
        np.cumsum(cnt_p, out=offsets[1:])
        # gather ids: concatenate groups in perm order
        src_start = np.repeat(starts[perm], cnt_p)
        within = np.arange(order.size) - np.repeat(offsets[:-1], cnt_p)
        gids = order[src_start + within] if order.size else order
        self.Ngather = int(perm.size)
        self.offsets = offsets.astype(np.int32)
        self.gather_ids = gids.astype(np.int32)
        deg = np.ones(self.N)
        deg[gids] = np.repeat(cnt_p, cnt_p)
        self.inv_degree = 1.0 / deg


def dirichlet_mask_ids(N: int, Nelements: int, EToB: np.ndarray, ogs_mesh: Ogs, orc):
    """maskIds of `ellipticOgs` (src/solvers/elliptic/ellipticOgs.cpp:4-75):
    node-wise min of the face BC flags, gs-min over shared nodes, ids whose flag
    is DIRICHLET(1) in ascending order."""
    Np = (N + 1) ** 3
    large = 1 << 20
    fn = face_nodes(N)
    mapB = np.full((Nelements, Np), large, dtype=np.int32)
    etob = np.asarray(EToB).reshape(Nelements, 6)
    for f in range(6):
        bc = etob[:, f]
        sel = bc > 0
        if sel.any():
            sub = mapB[np.ix_(np.flatnonzero(sel), fn[f])]
            mapB[np.ix_(np.flatnonzero(sel), fn[f])] = np.minimum(sub, bc[sel, None])
    mapB = np.ascontiguousarray(mapB.reshape(-1))
    orc.gs_min_i(ogs_mesh, mapB)
    return np.flatnonzero(mapB == 1).astype(np.int32), mapB

"""GPU parity of every hand-written kernel against the oracle (CPU restatement of the reference's
serial kernels), called through the C ABI.

Tolerances (BASELINE.json north_star): 1e-12 relative in fp64, 1e-5 in the fp32 multigrid kernels;
integer/index work (gather-scatter maps, masks) bit-exact.  Differences come only from FMA
contraction / summation order.
"""
import numpy as np
import pytest

from nekrs_b200 import lib, meshgen, ops
from nekrs_b200.lib import DeviceBuffer as DB
from oracle import sem

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def relerr(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    assert lib.device_count() > 0, "no CUDA device: the product has no CPU fallback"


# ------------------------------------------------------------------------------------ axhelm
@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6])
def test_axhelm_poisson(orc, N, dt, variant):
    E, Np = 37, (N + 1) ** 3
    r = rng(100 + N)
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g).astype(dt)
    ggeo = r.random((E, 7, Np)).astype(dt)
    q = r.random(E * Np).astype(dt)
    el = r.permutation(E)[: E - 5].astype(np.int32)
    ref = np.full(E * Np, -3.0, dtype=dt)
    orc.ax(N, el, ggeo, D, q, ref)
    d_Aq = DB(like=np.full(E * Np, -3.0, dtype=dt))
    lam0 = DB(like=np.ones(1, dtype=dt))
    ops.ellipticPartialAxCoeffHex3D(N, DB(like=el), DB(like=ggeo), D, DB(like=q), d_Aq, Nelements=el.size,
                                    lambda0=lam0, variant=variant, dtype=dt)
    out = d_Aq.download(dt)
    assert relerr(out, ref) < TOL[dt] * (10 if dt == np.float32 else 1)
    untouched = np.setdiff1d(np.arange(E), el)
    assert np.all(out.reshape(E, Np)[untouched] == -3.0)


@pytest.mark.parametrize("variant", [0, 1, 4, 5, 6, -1])  # 4-6, -1: TMA ring (constant coefficients), else pencil
@pytest.mark.parametrize("lambda_field", [False, True])
def test_axhelm_helmholtz(orc, variant, lambda_field):
    N, E = 7, 11
    Np = 512
    r = rng(7)
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    ggeo = r.random((E, 7, Np))
    q = r.random(E * Np)
    el = np.arange(E, dtype=np.int32)
    if lambda_field:
        lam0, lam1 = r.random(E * Np) + 0.5, r.random(E * Np)
    else:
        lam0, lam1 = np.array([1.3]), np.array([0.7])
    ref = np.zeros(E * Np)
    # oracle: p_lambda selects per-node coefficients
    import ctypes as C
    S = np.ascontiguousarray(D.T)
    orc.lib.orc_ax_d(C.c_int(E), C.c_int(0), C.c_int(0), el.ctypes.data_as(C.c_void_p),
                     ggeo.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p),
                     lam0.ctypes.data_as(C.c_void_p), lam1.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p),
                     ref.ctypes.data_as(C.c_void_p), C.c_int(N + 1), C.c_int(0), C.c_int(1 if lambda_field else 0))
    d_Aq = DB.zeros(E * Np, np.float64)
    ops.ellipticPartialAxCoeffHex3D(N, DB(like=el), DB(like=ggeo), D, DB(like=q), d_Aq, lambda0=DB(like=lam0),
                                    lambda1=DB(like=lam1), poisson=False, lambda_field=lambda_field, variant=variant)
    assert relerr(d_Aq.download(), ref) < 1e-12


@pytest.mark.parametrize("N,dt", [(5, np.float64), (9, np.float64), (9, np.float32), (11, np.float32)])
@pytest.mark.parametrize("poisson", [True, False])
def test_axhelm_tma_ring_other_orders(orc, N, dt, poisson):
    """The persistent TMA-ring kernel at the even Nq other than 8 (axhelm_tma_nq.cu; padded consumer groups): more
    elements than one round of the ring per group, a permuted element list, and the default dispatch (-1) must pick
    the same kernel as variant 4 (bit-identical output)."""
    E, Np = 1500, (N + 1) ** 3
    r = rng(300 + N)
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g).astype(dt)
    ggeo = r.random((E, 7, Np)).astype(dt)
    q = r.random(E * Np).astype(dt)
    el = r.permutation(E)[: E - 7].astype(np.int32)
    lam0, lam1 = np.full(1, 1.3, dtype=dt), np.full(1, 0.7, dtype=dt)
    ref = np.full(E * Np, -3.0, dtype=dt)
    orc.ax(N, el, ggeo, D, q, ref, lambda0=lam0, lambda1=lam1, poisson=poisson)
    outs = []
    for variant in (4, -1):
        d_Aq = DB(like=np.full(E * Np, -3.0, dtype=dt))
        ops.ellipticPartialAxCoeffHex3D(N, DB(like=el), DB(like=ggeo), D, DB(like=q), d_Aq, Nelements=el.size,
                                        lambda0=DB(like=lam0), lambda1=DB(like=lam1), poisson=poisson,
                                        variant=variant, dtype=dt)
        outs.append(d_Aq.download(dt))
    assert relerr(outs[0], ref) < TOL[dt] * (10 if dt == np.float32 else 1)
    assert np.array_equal(outs[0], outs[1])
    untouched = np.setdiff1d(np.arange(E), el)
    assert np.all(outs[0].reshape(E, Np)[untouched] == -3.0)


def test_axhelm_empty_list():
    D = np.eye(8)
    ops.ellipticPartialAxCoeffHex3D(7, DB(like=np.zeros(1, np.int32)), None, D, None, None, Nelements=0,
                                    dtype=np.float64)


# ------------------------------------------------------------------------------------ gather-scatter
@pytest.mark.parametrize("N,nel", [(1, (3, 2, 2)), (3, (3, 3, 2)), (7, (3, 2, 2))])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_ogs_maps_and_gather_scatter(orc, N, nel, dt):
    m = meshgen.box_mesh(N, nel)
    ids = m.global_ids.copy()
    ids[::17] = 0  # some masked nodes
    o_ref = sem.Ogs(ids)
    o = ops.Ogs(ids)
    off, gid = o.local_maps()
    assert o.NlocalGather == o_ref.Ngather
    assert np.array_equal(off, o_ref.offsets)          # bit-exact maps
    assert np.array_equal(gid, o_ref.gather_ids)
    assert np.array_equal(o.inv_degree(), o_ref.inv_degree)
    r = rng(5)
    k, stride = 2, ids.size + 24
    v = r.random(k * stride).astype(dt)
    ref = v.copy()
    orc.gs_add(o_ref, ref, k=k, stride=stride)
    d = DB(like=v)
    o.gather_scatter(d, k=k, stride=stride)
    assert np.array_equal(d.download(dt), ref)         # same summation order -> bit-exact
    # kernel-level CSR entry point
    d2 = DB(like=v)
    ops.gatherScatterMany_add(o_ref.Ngather, DB(like=o_ref.offsets), DB(like=o_ref.gather_ids), d2, k=k, stride=stride)
    assert np.array_equal(d2.download(dt), ref)


def test_gs_sanity_invariant():
    """gs(1) * invDegree sums to E*Np (meshParallelGatherScatterSetup.cpp:136-165)."""
    m = meshgen.box_mesh(5, (4, 3, 2))
    o = ops.Ogs(m.global_ids)
    n = m.global_ids.size
    d = DB(like=np.ones(n))
    o.gather_scatter(d)
    s = np.sum(d.download() * o.inv_degree())
    assert abs(s - n) / n < 1e-15


def test_gs_general_rows_and_mask():
    # rows of length 3, 5, 6 exercise the CSR bucket; mask ids zeroed in the same launch
    ids = np.array([1, 2, 3, 1, 1, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 0, 0, 9, 2], dtype=np.int64)
    o = ops.Ogs(ids)
    assert o.nGen == 3 and o.nPairs == 1
    v = np.arange(1.0, ids.size + 1)
    d = DB(like=v)
    mask_ids = np.array([16, 17], dtype=np.int32)
    o.gather_scatter(d, mask_ids=DB(like=mask_ids))
    ref = v.copy()
    for g in (1, 2, 4, 5):
        ref[ids == g] = v[ids == g].sum()
    ref[[16, 17]] = 0
    assert np.array_equal(d.download(), ref)


def test_mask(orc):
    q = rng(1).random(1000)
    ids = np.unique(rng(2).integers(0, 1000, 77)).astype(np.int32)
    d = DB(like=q)
    ops.mask(DB(like=ids), d)
    ref = q.copy()
    orc.mask(ids, ref)
    assert np.array_equal(d.download(), ref)


# ------------------------------------------------------------------------------------ linAlg
@pytest.mark.parametrize("N", [16, 255, 256, 4096 + 3, 300001])  # ethier/ci.inc:729-780 sizes and beyond
def test_linalg_fp64(orc, N):
    r = rng(N)
    x, y, w = r.random(N), r.random(N), r.random(N)
    dx, dy, dw = DB(like=x), DB(like=y), DB(like=w)
    yr = y.copy()
    orc.axpby(N, 0.3, x, -1.7, yr)
    ops.axpby(N, 0.3, dx, -1.7, dy)
    assert relerr(dy.download(), yr) < 1e-15
    zr = np.zeros(N)
    dz = DB.zeros(N, np.float64)
    orc.axmyz(N, 1.5, x, y, zr)
    ops.axmyz(N, 1.5, dx, DB(like=y), dz)
    assert np.array_equal(dz.download(), zr)
    ref = orc.weighted_inner_prod(N, w, x, y)
    got = ops.weightedInnerProdMany(N, 1, 0, dw, dx, DB(like=y))
    assert abs(got - ref) / abs(ref) < 1e-12
    ref = orc.weighted_norm2_sq(N, w, x)
    got = ops.weightedNorm2Many(N, 1, 0, dw, dx)
    assert abs(got - ref) / abs(ref) < 1e-12
    assert abs(ops.sum(N, dx) - orc.sum(N, x)) / N < 1e-13
    # determinism: same launch twice -> identical bits
    a = ops.weightedNorm2Many(N, 1, 0, dw, dx)
    assert a == got


def test_linalg_fp32_and_casts(orc):
    N = 10007
    r = rng(9)
    x, y = r.random(N).astype(np.float32), r.random(N).astype(np.float32)
    dx, dy = DB(like=x), DB(like=y)
    yr = y.copy()
    orc.axpby(N, 0.25, x, 2.0, yr)
    ops.axpby(N, 0.25, dx, 2.0, dy)
    assert relerr(dy.download(), yr) < 1e-6
    xd = r.random(N)
    f = DB.zeros(N, np.float32)
    ops.copyDfloatToPfloat(N, DB(like=xd), f)
    assert np.array_equal(f.download(), xd.astype(np.float32))
    d = DB.zeros(N, np.float64)
    ops.copyPfloatToDfloat(N, f, d)
    assert np.array_equal(d.download(), xd.astype(np.float32).astype(np.float64))
    ops.fill(N, 2.5, f)
    assert np.all(f.download() == np.float32(2.5))
    ops.scale(N, 2.0, f)
    assert np.all(f.download() == np.float32(5.0))


def test_update_pcg_fused(orc):
    N = 123457
    r = rng(3)
    w, Ap, p = r.random(N), r.random(N), r.random(N)
    rr, x = r.random(N), r.random(N)
    ref_r = rr.copy()
    ref = orc.update_pcg(N, w, Ap, 0.37, ref_r)
    ref_x = x.copy()
    orc.axpby(N, 0.37, p, 1.0, ref_x)
    dr, dx = DB(like=rr), DB(like=x)
    got = ops.ellipticBlockUpdatePCG(N, DB(like=w), DB(like=Ap), 0.37, dr, p=DB(like=p), x=dx)
    assert abs(got - ref) / ref < 1e-12
    assert relerr(dr.download(), ref_r) < 1e-15
    assert relerr(dx.download(), ref_x) < 1e-15


def test_gmres_kernels(orc):
    r = rng(4)
    N, off, m = 50001, 50176, 7
    w = r.random(N)
    V = r.random(off * m)
    y = r.random(m)
    wv = r.random(off)
    ref_w = wv.copy()
    ref = orc.gram_schmidt(N, off, m, w, y, V, ref_w)
    dwv = DB(like=wv)
    got = ops.gramSchmidtOrthogonalization(N, off, m, DB(like=w), DB(like=y), DB(like=V), dwv)
    assert abs(got - ref) / ref < 1e-12
    assert relerr(dwv.download()[:N], ref_w[:N]) < 1e-13
    x = r.random(off)
    ref_x = x.copy()
    orc.update_pgmres_solution(N, off, m, y, V, ref_x)
    dxx = DB(like=x)
    ops.updatePGMRESSolution(N, off, m, DB(like=y), DB(like=V), dxx)
    assert relerr(dxx.download()[:N], ref_x[:N]) < 1e-13
    b, Ax = r.random(N), r.random(N)
    ref_r = np.zeros(N)
    ref = orc.fused_residual_and_norm(N, w, b, Ax, ref_r)
    dres = DB.zeros(N, np.float64)
    got = ops.fusedResidualAndNorm(N, DB(like=w), DB(like=b), DB(like=Ax), dres)
    assert abs(got - ref) / ref < 1e-12
    assert np.array_equal(dres.download(), ref_r)
    X = r.random(off * m)
    refm = orc.weighted_inner_prod_multi(N, m, off, w, X, b)
    gotm = ops.weightedInnerProdMulti(N, m, off, DB(like=w), DB(like=X), DB(like=b))
    assert np.max(np.abs(gotm - refm) / np.abs(refm)) < 1e-12


def test_chebyshev_updates(orc):
    N = 40001
    r = rng(6)
    f = np.float32
    SAd, d, res, x = (r.random(N).astype(f) for _ in range(4))
    rd, rr, rx = d.copy(), res.copy(), x.copy()
    orc.update_chebyshev(N, 0.6, 1.2, SAd, rd, rr, rx)
    dd, dr, dx = DB(like=d), DB(like=res), DB(like=x)
    ops.updateChebyshev(N, 0.6, 1.2, DB(like=SAd), dd, dr, dx)
    assert relerr(dd.download(), rd) < 1e-6 and np.array_equal(dr.download(), rr) and np.array_equal(dx.download(), rx)
    rr, rx = res.copy(), x.copy()
    orc.update_fourth_chebyshev(N, 0.8, SAd, d, rr, rx)
    dr, dx = DB(like=res), DB(like=x)
    ops.updateFourthKindChebyshev(N, 0.8, DB(like=SAd), DB(like=d), dr, dx)
    assert relerr(dx.download(), rx) < 1e-6 and np.array_equal(dr.download(), rr)


# ------------------------------------------------------------------------------------ FDM / transfers
@pytest.mark.parametrize("N", [1, 2, 3, 5, 7, 9])
@pytest.mark.parametrize("restrict", [1, 0])
@pytest.mark.parametrize("fdm_variant", [0, 1, 2])
def test_fdm(orc, N, restrict, fdm_variant):
    import ctypes
    lib.call("nrsb_set_fdm_variant", ctypes.c_int(fdm_variant))
    E = 23
    Nq, Nqe = N + 1, N + 3
    r = rng(20 + N)
    f = np.float32
    u = r.random(E * Nq ** 3).astype(f)
    w_ref = np.zeros(E * Nqe ** 3, f)
    orc.pre_fdm(E, N, u, w_ref)
    d_w = DB.zeros(E * Nqe ** 3, f)
    ops.preFDM(N, E, DB(like=u), d_w)
    assert np.array_equal(d_w.download(), w_ref)
    w_in = (w_ref + r.random(w_ref.size).astype(f)).astype(f)   # non-trivial overlap planes
    Sx, Sy, Sz = (((r.random(E * Nqe * Nqe) - 0.5) * 0.5).astype(f) for _ in range(3))
    invL = r.random(E * Nqe ** 3).astype(f)
    wts = r.random(E * Nq ** 3).astype(f)
    nsu = E * (Nq ** 3 if restrict else Nqe ** 3)
    Su_ref = np.zeros(nsu, f)
    wr = w_in.copy()
    orc.fused_fdm(E, N, Su_ref, Sx, Sy, Sz, invL, wts, wr, restrict)
    d_Su, d_u = DB.zeros(nsu, f), DB(like=w_in)
    ops.fusedFDM(N, restrict, E, DB(like=np.arange(E, dtype=np.int32)), d_Su, DB(like=Sx), DB(like=Sy), DB(like=Sz),
                 DB(like=invL), DB(like=wts), d_u)
    Su = d_Su.download()
    assert relerr(Su, Su_ref) < 1e-5
    if not restrict:
        # u: only the six overlap planes (face interiors) are defined by the reference
        a = d_u.download().reshape(E, Nqe, Nqe, Nqe)
        b = wr.reshape(E, Nqe, Nqe, Nqe)
        s = slice(1, Nqe - 1)
        scale = np.max(np.abs(Su_ref))
        for pa, pb in ((a[:, 0, s, s], b[:, 0, s, s]), (a[:, -1, s, s], b[:, -1, s, s]),
                       (a[:, s, 0, s], b[:, s, 0, s]), (a[:, s, -1, s], b[:, s, -1, s]),
                       (a[:, s, s, 0], b[:, s, s, 0]), (a[:, s, s, -1], b[:, s, s, -1])):
            assert np.max(np.abs(pa - pb)) / scale < 1e-5
        out_ref = np.zeros(E * Nq ** 3, f)
        orc.post_fdm(E, N, wr, Su_ref, out_ref, wts)
        d_out = DB.zeros(E * Nq ** 3, f)
        ops.postFDM(N, E, DB(like=wr), DB(like=Su_ref), d_out, DB(like=wts))
        assert relerr(d_out.download(), out_ref) < 1e-6


@pytest.mark.parametrize("Nf,Nc", [(7, 3), (3, 1), (7, 5), (5, 3), (9, 5), (5, 1), (7, 1), (2, 1)])
def test_transfers(orc, Nf, Nc):
    E = 29
    r = rng(Nf * 10 + Nc)
    f = np.float32
    gf, _ = sem.jacobi_gll(Nf)
    gc, _ = sem.jacobi_gll(Nc)
    R = sem.interpolation_matrix_1d(gc, gf).T.copy().astype(f)
    qf = r.random(E * (Nf + 1) ** 3).astype(f)
    ref = np.zeros(E * (Nc + 1) ** 3, f)
    orc.coarsen(E, Nf, Nc, R, qf, ref)
    d = DB.zeros(ref.size, f)
    ops.ellipticPreconCoarsenHex3D(Nf, Nc, E, R, DB(like=qf), d)
    assert relerr(d.download(), ref) < 1e-6
    fine = r.random(qf.size).astype(f)
    pref = fine.copy()
    orc.prolongate(E, Nf, Nc, R, ref, pref)
    dfine = DB(like=fine)
    ops.ellipticPreconProlongateHex3D(Nf, Nc, E, R, DB(like=ref), dfine)
    assert relerr(dfine.download(), pref) < 1e-6


@pytest.mark.parametrize("N", [1, 3, 7])
def test_geometric_factors(orc, N):
    m = meshgen.box_mesh(N, (3, 2, 2), kershaw_eps=0.3)
    g, w = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    E = m.Nelements
    ref, Jref = orc.geometric_factors(E, N, D, w, m.x, m.y, m.z)
    d_g, d_J = DB.zeros(E * 7 * m.Np, np.float64), DB.zeros(E * m.Np, np.float64)
    ops.geometricFactorsHex3D(N, E, D, w, DB(like=m.x), DB(like=m.y), DB(like=m.z), d_g, d_J)
    assert relerr(d_g.download(), ref.ravel()) < 1e-12
    assert relerr(d_J.download(), Jref.ravel()) < 1e-12
    assert np.all(d_J.download() > 0)   # meshGeometricFactorsHex3D.cpp:52-58


# ------------------------------------------------------------------------------------ diagonal, linAlg "Many"
@pytest.mark.parametrize("N", [1, 3, 7, 9])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("poisson,field,Nfields", [(True, False, 1), (False, False, 1), (False, True, 1), (True, True, 3),
                                                   (False, True, 3)])
def test_build_diagonal(N, dt, poisson, field, Nfields):
    from oracle import kernels as K
    E, Np = 23, (N + 1) ** 3
    r = rng(50 + N)
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g).astype(dt)
    ggeo = (r.random((E, 7, Np)) + 0.1).astype(dt)
    offset = E * Np + 40
    loffset = E * Np + 8
    if field:
        lam0 = (r.random(Nfields * loffset) + 0.5).astype(dt)
        lam1 = r.random(Nfields * loffset).astype(dt)
    else:
        lam0, lam1 = np.array([1.3], dtype=dt), np.array([0.7], dtype=dt)
    ref = K.build_diagonal(N, E, ggeo, D, lam0, lam1, poisson=poisson, lambda_field=field, Nfields=Nfields,
                           offset=offset, loffset=loffset)
    d_Aq = DB.zeros(Nfields * offset, dt)
    ops.ellipticBlockBuildDiagonalHex3D(N, E, DB(like=ggeo), D, DB(like=lam0), DB(like=lam1), d_Aq, Nfields=Nfields,
                                        offset=offset, loffset=loffset, poisson=poisson, lambda_field=field, dtype=dt)
    out = d_Aq.download(dt)
    for l in range(Nfields):
        sl = slice(l * offset, l * offset + E * Np)
        assert relerr(out[sl], ref[sl]) < (1e-12 if dt == np.float64 else 2e-5)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_linalg_many_family(dt):
    """scaleMany, add, abs, axmy(Many), axmyzMany, adyMany, axdy, axpbyzMany (linAlg.hpp:70-142) against numpy."""
    prec = 8 if dt == np.float64 else 4
    N, Nf, off = 1000, 3, 1031
    r = rng(77)
    x = (r.random(Nf * off) + 0.5).astype(dt)
    y = (r.random(Nf * off) - 0.5).astype(dt)
    tol = 1e-14 if dt == np.float64 else 1e-6

    def fields(a):
        return np.concatenate([a[f * off:f * off + N] for f in range(Nf)])

    d = DB(like=x)
    ops.linalg_many("scaleMany", prec, N, Nf, off, 1.7, d)
    assert relerr(fields(d.download(dt)), fields(x) * dt(1.7)) < tol
    untouched = d.download(dt)[N:off]
    assert np.array_equal(untouched, x[N:off])          # the gap between fields is not written
    d = DB(like=y)
    ops.linalg_many("add", prec, N, 0.25, d)
    assert relerr(d.download(dt)[:N], y[:N] + dt(0.25)) < tol
    d = DB(like=y)
    ops.linalg_many("abs", prec, N, d)
    assert np.array_equal(d.download(dt)[:N], np.abs(y[:N]))
    d = DB(like=y)
    ops.linalg_many("axmy", prec, N, 2.0, DB(like=x), d)
    assert relerr(d.download(dt)[:N], dt(2) * x[:N] * y[:N]) < tol
    for mode in (0, 1):
        d = DB(like=y)
        ops.linalg_many("axmyMany", prec, N, Nf, off, mode, 0.5, DB(like=x), d)
        xx = fields(x) if mode else np.tile(x[:N], Nf)
        assert relerr(fields(d.download(dt)), dt(0.5) * xx * fields(y)) < tol
    d = DB.zeros(Nf * off, dt)
    ops.linalg_many("axmyzMany", prec, N, Nf, off, 3.0, DB(like=x), DB(like=y), d)
    assert relerr(fields(d.download(dt)), dt(3) * fields(x) * fields(y)) < tol
    d = DB(like=x)
    ops.linalg_many("adyMany", prec, N, Nf, off, 1.0, d)
    assert relerr(fields(d.download(dt)), dt(1) / fields(x)) < tol
    d = DB(like=x)
    ops.linalg_many("axdy", prec, N, 2.0, DB(like=y), d)
    assert relerr(d.download(dt)[:N], dt(2) * y[:N] / x[:N]) < tol
    d = DB.zeros(Nf * off, dt)
    ops.linalg_many("axpbyzMany", prec, N, Nf, off, 2.0, DB(like=x), -3.0, DB(like=y), d)
    assert relerr(fields(d.download(dt)), dt(2) * fields(x) + dt(-3) * fields(y)) < tol


@pytest.mark.parametrize("N", [3, 7, 9])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lambda_field", [False, True])
def test_block_axhelm(orc, N, dt, lambda_field):
    """ellipticBlockPartialAxCoeffHex3D: three fields sharing the geometric factors, per-field coefficients."""
    E, Np = 19, (N + 1) ** 3
    r = rng(60 + N)
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g).astype(dt)
    ggeo = r.random((E, 7, Np)).astype(dt)
    offset, loffset = E * Np + 24, E * Np + 8
    q = r.random(3 * offset).astype(dt)
    if lambda_field:
        lam0, lam1 = (r.random(3 * loffset) + 0.5).astype(dt), r.random(3 * loffset).astype(dt)
    else:
        lam0, lam1 = np.zeros(3 * loffset, dt), np.zeros(3 * loffset, dt)
        lam0[[0, loffset, 2 * loffset]] = [1.1, 1.2, 1.3]
        lam1[[0, loffset, 2 * loffset]] = [0.5, 0.6, 0.7]
    el = r.permutation(E)[: E - 3].astype(np.int32)
    ref = np.full(3 * offset, -3.0, dtype=dt)
    orc.ax_block(N, el, ggeo, D, q, ref, lam0, lam1, offset, loffset, lambda_field=lambda_field)
    d_Aq = DB(like=np.full(3 * offset, -3.0, dtype=dt))
    ops.ellipticBlockPartialAxCoeffHex3D(N, el.size, offset, loffset, DB(like=el), DB(like=ggeo), D, DB(like=lam0),
                                         DB(like=lam1), DB(like=q), d_Aq, lambda_field=lambda_field, dtype=dt)
    out = d_Aq.download(dt)
    assert relerr(out, ref) < TOL[dt] * (10 if dt == np.float32 else 1)
    untouched = np.setdiff1d(np.arange(E), el)
    for f in range(3):
        assert np.all(out[f * offset:f * offset + E * Np].reshape(E, Np)[untouched] == -3.0)


@pytest.mark.parametrize("N", [1, 3, 4, 7, 9, 11])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lambda_field", [False, True])
def test_stress_axhelm(orc, N, dt, lambda_field):
    """ellipticStressPartialAxCoeffHex3D: the coupled viscous-stress operator on three fields (the oracle is pinned
    bit-exact to the reference's kernel in tests/test_oracle_vs_ref.py)."""
    if N == 11 and dt == np.float64:
        E = 5
    else:
        E = 19
    Np = (N + 1) ** 3
    r = rng(80 + N)
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g).astype(dt)
    vgeo = (r.random((E, 12, Np)) - 0.3).astype(dt)
    offset, loffset = E * Np + 24, E * Np + 8
    q = r.random(3 * offset).astype(dt)
    if lambda_field:
        lam0, lam1 = (r.random(3 * loffset) + 0.5).astype(dt), r.random(3 * loffset).astype(dt)
    else:
        lam0, lam1 = np.zeros(3 * loffset, dt), np.zeros(3 * loffset, dt)
        lam0[[0, loffset, 2 * loffset]] = [1.1, 1.2, 1.3]
        lam1[[0, loffset, 2 * loffset]] = [0.5, 0.6, 0.7]
    el = r.permutation(E)[: E - 2].astype(np.int32)
    ref = np.full(3 * offset, -3.0, dtype=dt)
    orc.ax_stress(N, el, vgeo, D, q, ref, lam0, lam1, offset, loffset, lambda_field=lambda_field)
    d_Aq = DB(like=np.full(3 * offset, -3.0, dtype=dt))
    ops.ellipticStressPartialAxCoeffHex3D(N, el.size, offset, loffset, DB(like=el), DB(like=vgeo), D, DB(like=lam0),
                                          DB(like=lam1), DB(like=q), d_Aq, lambda_field=lambda_field, dtype=dt)
    out = d_Aq.download(dt)
    assert relerr(out, ref) < TOL[dt] * (10 if dt == np.float32 else 1)
    untouched = np.setdiff1d(np.arange(E), el)
    for f in range(3):
        assert np.all(out[f * offset:f * offset + E * Np].reshape(E, Np)[untouched] == -3.0)

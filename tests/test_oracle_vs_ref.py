"""Pin the C restatement (oracle/nrs_oracle.c) bit-for-bit against the reference's
own SERIAL kernels compiled from /root/reference (oracle/_ref/*.so).

CPU only.  Skips (does not fail) when oracle/_ref has not been built, e.g. in a
checkout without /root/reference and without the prebuilt libraries.
"""
import numpy as np
import pytest

from oracle import kernels as K
from oracle import sem

needs = lambda name: pytest.mark.skipif(not K.ref_available(name), reason="oracle/_ref/%s.so not built" % name)


def rng(seed=1234):
    return np.random.Generator(np.random.PCG64(seed))


@pytest.mark.parametrize("N", [1, 2, 3, 5, 7, 9])
@pytest.mark.parametrize("prec", ["d", "f"])
def test_ax_bit_exact(orc, N, prec):
    if not K.ref_available("ax_%s_N%d_poisson" % (prec, N)):
        pytest.skip("ref not built")
    dt = np.float64 if prec == "d" else np.float32
    E, Np = 7, (N + 1) ** 3
    r = rng(N)
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g).astype(dt)
    ggeo = r.random((E, 7, Np)).astype(dt)
    q = r.random(E * Np).astype(dt)
    el = np.array([3, 0, 6, 5, 1], dtype=np.int32)  # partial, permuted list
    a = np.full(E * Np, -7.0, dtype=dt)
    b = a.copy()
    K.RefAx(N, prec)(el, ggeo, D, q, a)
    orc.ax(N, el, ggeo, D, q, b)
    assert np.array_equal(a, b)
    assert np.all(a.reshape(E, Np)[[2, 4]] == -7.0)  # unlisted elements untouched


@needs("ax_d_N7_helmholtz")
def test_ax_helmholtz_bit_exact(orc):
    N, E = 7, 3
    Np = 512
    r = rng(5)
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    ggeo = r.random((E, 7, Np))
    q = r.random(E * Np)
    el = np.arange(E, dtype=np.int32)
    lam0, lam1 = np.array([1.3]), np.array([0.7])
    a, b = np.zeros(E * Np), np.zeros(E * Np)
    K.RefAx(N, "d", poisson=False)(el, ggeo, D, q, a, lam0, lam1)
    orc.ax(N, el, ggeo, D, q, b, lam0, lam1, poisson=False)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("N", [1, 3, 7])
@pytest.mark.parametrize("restrict", [1, 0])
def test_fdm_bit_exact(orc, N, restrict):
    if not K.ref_available("fdm_N%d_r%d" % (N, restrict)):
        pytest.skip("ref not built")
    E = 5
    Nq, Nqe = N + 1, N + 3
    r = rng(10 + N)
    f32 = np.float32
    u = r.random(E * Nq ** 3).astype(f32)
    w1a, w1b = np.zeros(E * Nqe ** 3, f32), np.zeros(E * Nqe ** 3, f32)
    ref = K.RefFdm(N, restrict)
    ref.pre(E, u, w1a)
    orc.pre_fdm(E, N, u, w1b)
    assert np.array_equal(w1a, w1b)
    # perturb the overlap so the subtract step is exercised
    w1a += r.random(w1a.size).astype(f32)
    w1b[:] = w1a
    Sx, Sy, Sz = (r.random(E * Nqe * Nqe).astype(f32) - 0.5 for _ in range(3))
    invL = r.random(E * Nqe ** 3).astype(f32)
    wts = r.random(E * Nq ** 3).astype(f32)
    nsu = E * (Nq ** 3 if restrict else Nqe ** 3)
    Sua, Sub = np.zeros(nsu, f32), np.zeros(nsu, f32)
    ref.fused(E, Sua, Sx, Sy, Sz, invL, wts, w1a)
    orc.fused_fdm(E, N, Sub, Sx, Sy, Sz, invL, wts, w1b, restrict)
    assert np.array_equal(Sua, Sub)
    assert np.array_equal(w1a, w1b)
    if not restrict:
        oa, ob = np.zeros(E * Nq ** 3, f32), np.zeros(E * Nq ** 3, f32)
        ref.post(E, w1a, Sua, oa, wts)
        orc.post_fdm(E, N, w1b, Sub, ob, wts)
        assert np.array_equal(oa, ob)


@pytest.mark.parametrize("Nf,Nc", [(7, 3), (3, 1), (7, 5), (5, 3), (9, 5), (5, 1)])
def test_transfer_bit_exact(orc, Nf, Nc):
    if not K.ref_available("transfer_Nf%d_Nc%d" % (Nf, Nc)):
        pytest.skip("ref not built")
    E = 6
    r = rng(Nf * 10 + Nc)
    f32 = np.float32
    gf, _ = sem.jacobi_gll(Nf)
    gc, _ = sem.jacobi_gll(Nc)
    R = sem.interpolation_matrix_1d(gc, gf).T.copy().astype(f32)  # [NqC][NqF]
    qf = r.random(E * (Nf + 1) ** 3).astype(f32)
    a, b = np.zeros(E * (Nc + 1) ** 3, f32), np.zeros(E * (Nc + 1) ** 3, f32)
    ref = K.RefTransfer(Nf, Nc)
    ref.coarsen(E, R, qf, a)
    orc.coarsen(E, Nf, Nc, R, qf, b)
    assert np.array_equal(a, b)
    pa = r.random(E * (Nf + 1) ** 3).astype(f32)
    pb = pa.copy()
    ref.prolongate(E, R, a, pa)
    orc.prolongate(E, Nf, Nc, R, b, pb)
    assert np.array_equal(pa, pb)


@pytest.mark.parametrize("prec", ["d", "f"])
def test_linalg_bit_exact(orc, prec):
    if not K.ref_available("linalg_" + prec):
        pytest.skip("ref not built")
    dt = np.float64 if prec == "d" else np.float32
    ref = K.RefLinAlg(prec)
    r = rng(77)
    for N in (1, 16, 255, 256, 4097):  # BLOCKSIZE/16 .. 16*BLOCKSIZE-ish, as ethier/ci.inc:729-780
        x, y, w = (r.random(N).astype(dt) for _ in range(3))
        ya, yb = y.copy(), y.copy()
        ref.axpby_many(N, 0.3, x, -1.7, ya)
        orc.axpby(N, 0.3, x, -1.7, yb)
        assert np.array_equal(ya, yb)
        za, zb = np.zeros(N, dt), np.zeros(N, dt)
        ref.axmyz(N, 1.5, x, y, za)
        orc.axmyz(N, 1.5, x, y, zb)
        assert np.array_equal(za, zb)
        ya, yb = y.copy(), y.copy()
        ref.axmy(N, 0.9, x, ya)
        orc.axmy(N, 0.9, x, yb)
        assert np.array_equal(ya, yb)
        assert ref.weighted_inner_prod_many(N, w, x, y) == orc.weighted_inner_prod(N, w, x, y)
        assert ref.weighted_norm2_many(N, w, x) == orc.weighted_norm2_sq(N, w, x)


@needs("linalg_d")
def test_krylov_helpers_bit_exact(orc):
    ref = K.RefLinAlg("d")
    r = rng(3)
    N, off, m = 1000, 1024, 5
    w, Ap = r.random(N), r.random(N)
    ra, rb = r.random(N), None
    rb = ra.copy()
    assert ref.update_pcg(N, w, Ap, 0.37, ra) == orc.update_pcg(N, w, Ap, 0.37, rb)
    assert np.array_equal(ra, rb)
    V = r.random(off * m)
    y = r.random(m)
    wa = r.random(off)
    wb = wa.copy()
    assert ref.gram_schmidt(N, off, m, w, y, V, wa) == orc.gram_schmidt(N, off, m, w, y, V, wb)
    assert np.array_equal(wa, wb)
    xa = r.random(off)
    xb = xa.copy()
    ref.update_pgmres_solution(N, off, m, y, V, xa)
    orc.update_pgmres_solution(N, off, m, y, V, xb)
    assert np.array_equal(xa, xb)
    b, Ax = r.random(N), r.random(N)
    r1, r2 = np.zeros(N), np.zeros(N)
    assert ref.fused_residual_and_norm(N, w, b, Ax, r1) == orc.fused_residual_and_norm(N, w, b, Ax, r2)
    assert np.array_equal(r1, r2)
    xd = r.random(N)
    f1, f2 = np.zeros(N, np.float32), np.zeros(N, np.float32)
    ref.copy_d2f(xd, f1)
    orc.copy_d2f(xd, f2)
    assert np.array_equal(f1, f2)
    d1, d2 = np.zeros(N), np.zeros(N)
    ref.copy_f2d(f1, d1)
    orc.copy_f2d(f2, d2)
    assert np.array_equal(d1, d2)


@pytest.mark.parametrize("N", [3, 7])
@pytest.mark.parametrize("lambda_field", [False, True])
def test_block_ax_oracle_equals_reference_bit_exact(orc, N, lambda_field):
    """The three-field Helmholtz operator (ellipticBlockPartialAxCoeffHex3D.c): the oracle applies its scalar restatement
    field by field; bit-identical to the reference's block kernel compiled in place."""
    from oracle import kernels as K
    name = "axblock_d_N%d_lambda%d" % (N, 1 if lambda_field else 0)
    if not K.ref_available(name):
        pytest.skip("reference kernels not built here")
    E, Np = 7, (N + 1) ** 3
    r = np.random.Generator(np.random.PCG64(40 + N))
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    ggeo = r.random((E, 7, Np))
    offset, loffset = E * Np + 16, E * Np + 8
    q = r.random(3 * offset)
    if lambda_field:
        lam0, lam1 = r.random(3 * loffset) + 0.5, r.random(3 * loffset)
    else:
        lam0, lam1 = np.zeros(3 * loffset), np.zeros(3 * loffset)
        lam0[[0, loffset, 2 * loffset]] = [1.1, 1.2, 1.3]
        lam1[[0, loffset, 2 * loffset]] = [0.5, 0.6, 0.7]
    el = r.permutation(E).astype(np.int32)[:5]
    a = np.full(3 * offset, -2.0)
    b = a.copy()
    orc.ax_block(N, el, ggeo, D, q, a, lam0, lam1, offset, loffset, lambda_field=lambda_field)
    K.RefAxBlock(N, lambda_field)(el, ggeo, D, q, b, lam0, lam1, offset, loffset)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("N", [3, 7])
@pytest.mark.parametrize("lambda_field", [False, True])
def test_stress_ax_oracle_equals_reference_bit_exact(orc, N, lambda_field):
    """The coupled three-field stress operator (ellipticStressPartialAxCoeffHex3D.c): the oracle's three-pass
    restatement keeps every accumulation in the reference's order; bit-identical to the reference kernel compiled in
    place."""
    from oracle import kernels as K
    name = "axstress_d_N%d_lambda%d" % (N, 1 if lambda_field else 0)
    if not K.ref_available(name):
        pytest.skip("reference kernels not built here")
    E, Np = 7, (N + 1) ** 3
    r = np.random.Generator(np.random.PCG64(60 + N))
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    vgeo = r.random((E, 12, Np)) - 0.3
    offset, loffset = E * Np + 16, E * Np + 8
    q = r.random(3 * offset)
    if lambda_field:
        lam0, lam1 = r.random(3 * loffset) + 0.5, r.random(3 * loffset)
    else:
        lam0, lam1 = np.zeros(3 * loffset), np.zeros(3 * loffset)
        lam0[[0, loffset, 2 * loffset]] = [1.1, 1.2, 1.3]
        lam1[[0, loffset, 2 * loffset]] = [0.5, 0.6, 0.7]
    el = r.permutation(E).astype(np.int32)[:5]
    a = np.full(3 * offset, -2.0)
    b = a.copy()
    orc.ax_stress(N, el, vgeo, D, q, a, lam0, lam1, offset, loffset, lambda_field=lambda_field)
    K.RefAxStress(N, lambda_field)(el, vgeo, D, q, b, lam0, lam1, offset, loffset)
    assert np.array_equal(a, b)
    untouched = np.setdiff1d(np.arange(E), el)
    assert np.all(a[:E * Np].reshape(E, Np)[untouched] == -2.0)


def test_stress_ax_symmetric_and_rigid_motions(orc):
    """Properties of the stress form on a real mesh (vgeo from the trilinear map of a kershaw box, computed here):
    the element operator is symmetric, and rigid translations have zero stress (lambda1 = 0)."""
    from nekrs_b200 import meshgen
    N = 3
    m = meshgen.box_mesh(N, (2, 2, 1), kershaw_eps=0.5)
    E, Np, Nq = m.Nelements, m.Np, N + 1
    g, w = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    vgeo = np.zeros((E, 12, Np))
    X = [np.asarray(c).reshape(E, Nq, Nq, Nq) for c in (m.x, m.y, m.z)]
    dr = lambda f: np.einsum("im,ekjm->ekji", D, f)
    ds = lambda f: np.einsum("jm,ekmi->ekji", D, f)
    dtt = lambda f: np.einsum("km,emji->ekji", D, f)
    J = np.empty((E, Nq, Nq, Nq, 3, 3))
    for a, f in enumerate(X):
        J[..., a, 0], J[..., a, 1], J[..., a, 2] = dr(f), ds(f), dtt(f)
    Ji = np.linalg.inv(J)            # rows r,s,t ; columns x,y,z
    det = np.linalg.det(J)
    for a in range(3):
        for b in range(3):
            vgeo[:, 3 * a + b] = Ji[..., a, b].reshape(E, Np)
    W = (w[:, None, None] * w[None, :, None] * w[None, None, :]).reshape(1, Np)
    vgeo[:, 9] = det.reshape(E, Np)
    vgeo[:, 10] = det.reshape(E, Np) * W
    vgeo[:, 11] = 1.0 / vgeo[:, 10]
    offset = loffset = E * Np
    el = np.arange(E, dtype=np.int32)
    lam0 = np.zeros(3 * loffset)
    lam1 = np.zeros(3 * loffset)
    lam0[[0, loffset, 2 * loffset]] = 1.0
    r = np.random.Generator(np.random.PCG64(5))
    x1, x2 = r.random(3 * offset), r.random(3 * offset)
    A1, A2 = np.zeros(3 * offset), np.zeros(3 * offset)
    orc.ax_stress(N, el, vgeo, D, x1, A1, lam0, lam1, offset, loffset)
    orc.ax_stress(N, el, vgeo, D, x2, A2, lam0, lam1, offset, loffset)
    assert abs(np.dot(x2, A1) - np.dot(x1, A2)) < 1e-11 * abs(np.dot(x2, A1))
    t = np.concatenate([np.full(offset, 0.3), np.full(offset, -1.2), np.full(offset, 2.0)])
    At = np.zeros(3 * offset)
    orc.ax_stress(N, el, vgeo, D, t, At, lam0, lam1, offset, loffset)
    assert np.max(np.abs(At)) < 1e-11 * np.max(np.abs(A1))
    # rigid rotation about z: u = -y, v = x, w = 0  ->  symmetric gradient vanishes
    rot = np.concatenate([-np.asarray(m.y).ravel(), np.asarray(m.x).ravel(), np.zeros(offset)])
    Ar = np.zeros(3 * offset)
    orc.ax_stress(N, el, vgeo, D, rot, Ar, lam0, lam1, offset, loffset)
    assert np.max(np.abs(Ar)) < 1e-10 * np.max(np.abs(A1))


@pytest.mark.parametrize("stress", [False, True])
def test_oracle_block_solver(orc, stress):
    """The oracle's block solver (OBlockSolver: per-field masks over the unmasked numbering, block / stress operator,
    Jacobi-PCG): the converged solution satisfies the masked, assembled system, and the per-field Jacobi diagonal is the
    probed diagonal of the assembled block operator."""
    from nekrs_b200 import meshgen
    from oracle import driver
    mesh = meshgen.box_mesh(3, (2, 2, 2), kershaw_eps=0.3)
    E, Np = mesh.Nelements, mesh.Np
    base = np.asarray(mesh.EToB, dtype=np.int32).reshape(E, 6)
    etob = np.stack([base, base, base]).copy()
    etob[2][:, 5][etob[2][:, 5] > 0] = 4
    lam0, lam1 = [1.0, 1.3, 0.8], [0.6, 0.5, 0.9]
    opts = {"SOLVER": "PCG", "PRECONDITIONER": "JACOBI", "MAXIMUM ITERATIONS": "300", "SOLVER TOLERANCE": "1e-11"}
    s = driver.OBlockSolver(mesh, opts, etob.reshape(-1), lam0, lam1, orc, stress_form=stress)
    off, n = s.fieldOffset, E * Np
    r = np.random.Generator(np.random.PCG64(3))
    rhs = np.zeros(3 * off)
    for f in range(3):
        rhs[f * off:f * off + n] = r.random(n)
    x = s.solve(rhs, np.zeros(3 * off))
    Ax = np.zeros(3 * off)
    s.ell.operator(x, Ax)
    b = rhs.copy()
    s.ell.apply_mask(b)
    s.ell.gs(b)
    assert 0 < s.Niter < 300
    assert np.max(np.abs(Ax - b)) < 1e-9 * np.max(np.abs(b))
    assert np.all(x[s.ell.mask_ids] == 0.0)
    if not stress:   # block form: the Jacobi diagonal is exactly diag(A) (the stress form reuses it as an approximation)
        for node in (5, n // 2, off + 17, 2 * off + n - 3):
            e_ = np.zeros(3 * off)
            e_[node] = 1.0
            col = np.zeros(3 * off)
            s.ell.ax(e_, col)
            s.ell.gs(col)
            # assembled diagonal entry of the global node = sum over its copies of the local diagonal entries
            f, loc = divmod(node, off)
            copies = np.flatnonzero(mesh.global_ids == mesh.global_ids[loc]) + f * off
            ee = np.zeros(3 * off)
            ee[copies] = 1.0
            cc = np.zeros(3 * off)
            s.ell.ax(ee, cc)
            s.ell.gs(cc)
            assert abs(cc[node] - 1.0 / s.inv_diag[node]) < 1e-10 * abs(cc[node])

"""CPU checks of the oracle's setup restatement (numpy) and of the mesh generator: the
reference's own invariants (SURVEY.md §4 "self-checks inside the library") and analytic answers."""
import numpy as np
import pytest

from nekrs_b200 import meshgen
from oracle import sem


@pytest.mark.parametrize("N", [1, 2, 3, 5, 7, 9, 11])
def test_gll_and_dmatrix(N):
    x, w = sem.jacobi_gll(N)
    assert abs(w.sum() - 2.0) < 1e-14 and np.allclose(x, -x[::-1], atol=1e-15)
    # GLL integrates degree 2N-1 exactly
    for p in range(0, 2 * N, 2):
        assert abs(np.dot(w, x ** p) - 2.0 / (p + 1)) < 1e-13
    D = sem.dmatrix_1d(x)
    for p in range(N + 1):
        d = p * x ** (p - 1) if p else np.zeros_like(x)
        assert np.max(np.abs(D @ x ** p - d)) < 1e-11
    assert np.allclose(meshgen.gll_nodes(N), x, atol=1e-15)


def test_interpolation_matrix():
    xf, _ = sem.jacobi_gll(7)
    xc, _ = sem.jacobi_gll(3)
    P = sem.interpolation_matrix_1d(xc, xf)   # coarse -> fine
    for p in range(4):
        assert np.max(np.abs(P @ xc ** p - xf ** p)) < 1e-14
    assert np.allclose(P.sum(axis=1), 1.0)


def test_gs_test_c_analytic(orc):
    """gslib tests/gs_test.c:73-103: a 1-D chain of np 'ranks' each holding ids {r, r+1}: after
    gs(add) of ones the answers are 2,3,...,3,2 when every rank also holds id 1... restated on one
    rank: node j of segment r has id r+j+1; interior ids have 2 copies, the ends 1."""
    nseg = 5
    ids = np.array([[r + 1, r + 2] for r in range(nseg)], dtype=np.int64).ravel()
    o = sem.Ogs(ids)
    v = np.ones(ids.size)
    orc.gs_add(o, v)
    expect = np.array([1] + [2] * (2 * nseg - 2) + [1], dtype=float)
    assert np.array_equal(v, expect)
    assert np.array_equal(o.inv_degree, 1.0 / expect)


@pytest.mark.parametrize("N,nel", [(1, (2, 2, 2)), (3, (3, 2, 2)), (7, (2, 2, 3))])
def test_ogs_invariants(orc, N, nel):
    m = meshgen.box_mesh(N, nel)
    o = sem.Ogs(m.global_ids)
    nx, ny, nz = nel
    n_unique = (nx * N + 1) * (ny * N + 1) * (nz * N + 1)
    assert o.Ngather == n_unique
    assert o.offsets[-1] == m.global_ids.size
    # rows ordered by first local index, ascending inside a row (ogsSetup.cpp:196-249)
    first = o.gather_ids[o.offsets[:-1]]
    assert np.all(np.diff(first) > 0)
    for g in range(0, o.Ngather, max(1, o.Ngather // 50)):
        row = o.gather_ids[o.offsets[g]:o.offsets[g + 1]]
        assert np.all(np.diff(row) > 0)
        assert np.all(m.global_ids[row] == m.global_ids[row[0]])
    # gs(1)*invDegree sums to E*Np within 1e-15 (meshParallelGatherScatterSetup.cpp:136-165)
    v = np.ones(m.global_ids.size)
    orc.gs_add(o, v)
    assert abs(np.sum(v * o.inv_degree) - v.size) / v.size < 1e-15
    # shared copies agree in position
    for arr in (m.x, m.y, m.z):
        a = arr.copy()
        orc.gs_add(o, a)
        assert np.max(np.abs(a * o.inv_degree - arr)) < 1e-14


def test_geometric_factors_numpy_vs_c(orc):
    N = 5
    m = meshgen.box_mesh(N, (2, 3, 2), kershaw_eps=0.3)
    g, w = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    a, Ja = sem.geometric_factors(m.x, m.y, m.z, N)
    b, Jb = orc.geometric_factors(m.Nelements, N, D, w, m.x, m.y, m.z)
    assert np.max(np.abs(a - b)) / np.max(np.abs(b)) < 1e-13
    assert np.all(Jb > 0)
    # volume of the kershaw box is 1
    assert abs(b[:, 6].sum() - 1.0) < 1e-12


def test_ax_vs_independent_einsum(orc):
    """oracle Ax (restated serial kernel) against an independent dense formulation."""
    N, E = 4, 3
    Nq, Np = N + 1, (N + 1) ** 3
    r = np.random.Generator(np.random.PCG64(11))
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    ggeo = r.random((E, 7, Np))
    q = r.random(E * Np)
    out = np.zeros(E * Np)
    orc.ax(N, np.arange(E, dtype=np.int32), ggeo, D, q, out)
    Q = q.reshape(E, Nq, Nq, Nq)
    G = ggeo.reshape(E, 7, Nq, Nq, Nq)
    qr = np.einsum("im,ekjm->ekji", D, Q)
    qs = np.einsum("jm,ekmi->ekji", D, Q)
    qt = np.einsum("km,emji->ekji", D, Q)
    Gr = G[:, 0] * qr + G[:, 1] * qs + G[:, 4] * qt
    Gs = G[:, 1] * qr + G[:, 2] * qs + G[:, 3] * qt
    Gt = G[:, 4] * qr + G[:, 3] * qs + G[:, 5] * qt
    A = (np.einsum("mi,ekjm->ekji", D, Gr) + np.einsum("mj,ekmi->ekji", D, Gs) + np.einsum("mk,emji->ekji", D, Gt))
    assert np.max(np.abs(A.ravel() - out)) / np.max(np.abs(out)) < 1e-14


def test_kershaw_mesh_properties():
    m = meshgen.box_mesh(3, (6, 6, 6), kershaw_eps=0.3)
    assert m.x.min() == -0.5 and m.x.max() == 0.5 and abs(m.y.min() + 0.5) < 1e-15 and abs(m.z.max() - 0.5) < 1e-15
    _, J = sem.geometric_factors(m.x, m.y, m.z, 3)
    assert np.all(J > 0)
    # eps = 1 is the identity map
    a = meshgen.box_mesh(3, (6, 6, 6), kershaw_eps=1.0)
    b = meshgen.box_mesh(3, (6, 6, 6))
    assert np.allclose(a.y, b.y, atol=1e-15) and np.allclose(a.z, b.z, atol=1e-15)
    # all-Dirichlet boundary flags: 6*36 boundary faces
    assert (m.EToB == meshgen.DIRICHLET).sum() == 6 * 36


def test_partition_covers_mesh():
    N, nel = 2, (4, 4, 2)
    whole = meshgen.box_mesh(N, nel)
    ids = []
    ne = 0
    for r in range(4):
        p = meshgen.box_mesh(N, nel, rank=r, nranks=4)
        ne += p.Nelements
        ids.append(p.global_ids)
    assert ne == whole.Nelements
    assert np.array_equal(np.unique(np.concatenate(ids)), np.unique(whole.global_ids))


def test_helmholtz_branch_of_the_driver(orc):
    """oracle/driver.py with p_poisson = 0: the diagonal (ellipticBlockBuildDiagonalHex3D.okl: lambda0 * stiffness
    diagonal + lambda1 * GwJ) equals the operator applied to unit vectors, and Jacobi-PCG reproduces a manufactured
    solution.  (The product's Helmholtz handle is checked on the GPU in tests/test_gpu_elliptic.py.)"""
    from nekrs_b200 import meshgen
    from oracle import driver
    lam0, lam1 = 1.3, 0.7
    # one deformed element: no shared nodes, so gather-scatter is the identity and 1/invDiag is the local diagonal
    m1 = meshgen.box_mesh(2, (1, 1, 1), kershaw_eps=0.4)
    s1 = driver.OSolver(m1, {"SOLVER": "PCG", "PRECONDITIONER": "JACOBI"}, orc, poisson=False, lambda0=lam0, lambda1=lam1)
    n1 = m1.Nelements * m1.Np
    diag = np.zeros(n1)
    for i in range(n1):
        e_i, col = np.zeros(n1), np.zeros(n1)
        e_i[i] = 1.0
        s1.ell.ax(e_i, col)
        diag[i] = col[i]
    assert np.max(np.abs(1.0 / s1.inv_diag - diag) / diag) < 1e-13
    # manufactured solution on a 3x2x2 kershaw mesh
    mesh = meshgen.box_mesh(3, (3, 2, 2), kershaw_eps=0.4)
    opts = {"SOLVER": "PCG", "PRECONDITIONER": "JACOBI", "MAXIMUM ITERATIONS": "400", "SOLVER TOLERANCE": "1e-10",
            "LINEAR SOLVER STOPPING CRITERION": "RELATIVE"}
    s = driver.OSolver(mesh, opts, orc, poisson=False, lambda0=lam0, lambda1=lam1)
    x_true = np.sin(np.pi * mesh.x.ravel()) * np.sin(np.pi * mesh.y.ravel()) * np.sin(np.pi * mesh.z.ravel())
    x_true[s.ell.mask_ids] = 0.0
    b = np.zeros(x_true.size)
    s.ell.ax(x_true, b)                       # unassembled right-hand side, as ellipticSolve expects
    x = s.solve(b, np.zeros(x_true.size))
    assert 0 < s.Niter < 400
    assert np.max(np.abs(x - x_true)) / np.max(np.abs(x_true)) < 1e-7
    assert s.ell.allNeumann == 0


def test_diagonal_restatement_equals_unit_vector_probes_of_the_oracle_ax(orc):
    """oracle/kernels.build_diagonal (restated from ellipticBlockBuildDiagonalHex3D.okl, which has no serial .c) must
    equal diag(A_e) obtained by applying the pinned oracle Ax (variable coefficients, Helmholtz) to unit vectors."""
    import ctypes as C
    from oracle import kernels as K
    N, E = 2, 2
    Np = (N + 1) ** 3
    r = np.random.default_rng(1)
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    ggeo = r.random((E, 7, Np)) + 0.1
    lam0, lam1 = r.random(E * Np) + 0.5, r.random(E * Np)
    el = np.arange(E, dtype=np.int32)
    S = np.ascontiguousarray(D.T)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    diag = np.zeros(E * Np)
    for n in range(E * Np):
        q = np.zeros(E * Np)
        q[n] = 1
        out = np.zeros(E * Np)
        orc.lib.orc_ax_d(C.c_int(E), C.c_int(0), C.c_int(0), p(el), p(ggeo), p(D), p(S), p(lam0), p(lam1), p(q), p(out),
                         C.c_int(N + 1), C.c_int(0), C.c_int(1))
        diag[n] = out[n]
    ref = K.build_diagonal(N, E, ggeo, D, lam0, lam1, poisson=False, lambda_field=True)
    assert np.max(np.abs(ref - diag)) / np.max(np.abs(diag)) < 1e-14


def test_volume_factors_consistent_with_geometric_factors():
    """mesh->vgeo of the oracle (oracle/driver.py::volume_factors, the input of the stress-form operator) against
    the oracle's ggeo, which is pinned to the reference's geometricFactorsHex3D: G_ab = JW (grad a . grad b),
    GWJ = JW (meshGeometricFactorsHex3D.cpp; ids mesh3D.h:82-102)."""
    from nekrs_b200 import meshgen
    from oracle import driver
    from oracle.kernels import Orc
    orc = Orc()
    hm = meshgen.box_mesh(4, (2, 3, 2), kershaw_eps=0.3)
    m = driver.OMesh(orc, hm.N, hm.Nelements, hm.x, hm.y, hm.z, hm.global_ids, hm.EToB)
    v = driver.volume_factors(m)
    g = m.ggeo.reshape(m.E, 7, m.Np)
    r, s, t = v[:, 0:3], v[:, 3:6], v[:, 6:9]
    JW = v[:, 10]
    dot = lambda a, b: (a * b).sum(axis=1)
    ref = {0: dot(r, r), 1: dot(r, s), 2: dot(s, s), 3: dot(s, t), 4: dot(r, t), 5: dot(t, t)}
    scale = np.max(np.abs(g[:, :6]))
    for k, val in ref.items():
        assert np.max(np.abs(JW * val - g[:, k])) < 1e-12 * scale, k
    assert np.max(np.abs(JW - g[:, 6])) < 1e-13 * np.max(np.abs(g[:, 6]))
    assert np.all(v[:, 9] > 0) and np.allclose(v[:, 11] * JW, 1.0, rtol=1e-14)
    assert abs(JW.sum() - 1.0) < 1e-12      # kershaw map of the unit box: volume 1

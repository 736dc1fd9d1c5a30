"""Seeded inputs of the golden-vector cases (shared by make_golden.py, the CPU test of the oracle and the GPU
test of the CUDA kernels).  Every case is small: the whole fixture file is < 100 kB."""
import numpy as np

from oracle import sem


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def ax_case(N, prec, poisson=True, seed=None):
    dt = np.float64 if prec == "d" else np.float32
    E, Np = 4, (N + 1) ** 3
    r = rng(9000 + 10 * N + (0 if prec == "d" else 1) + (0 if poisson else 5) if seed is None else seed)
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g).astype(dt)
    ggeo = r.random((E, 7, Np)).astype(dt)
    q = r.random(E * Np).astype(dt)
    el = np.array([2, 0, 3], dtype=np.int32)  # partial, permuted list: element 1 must stay untouched
    lam0 = np.array([1.0 if poisson else 1.3], dtype=dt)
    lam1 = np.array([0.0 if poisson else 0.7], dtype=dt)
    return dict(N=N, dt=dt, E=E, Np=Np, D=D, ggeo=ggeo, q=q, el=el, lam0=lam0, lam1=lam1, poisson=poisson)


def fdm_case(N, restrict):
    E = 3
    Nq, Nqe = N + 1, N + 3
    r = rng(9100 + 10 * N + restrict)
    f32 = np.float32
    u = r.random(E * Nq ** 3).astype(f32)
    noise = r.random(E * Nqe ** 3).astype(f32)
    Sx, Sy, Sz = (r.random(E * Nqe * Nqe).astype(f32) - f32(0.5) for _ in range(3))
    invL = r.random(E * Nqe ** 3).astype(f32)
    wts = r.random(E * Nq ** 3).astype(f32)
    return dict(N=N, E=E, Nq=Nq, Nqe=Nqe, restrict=restrict, u=u, noise=noise, Sx=Sx, Sy=Sy, Sz=Sz, invL=invL, wts=wts)


def transfer_case(Nf, Nc):
    E = 3
    r = rng(9200 + 10 * Nf + Nc)
    f32 = np.float32
    gf, _ = sem.jacobi_gll(Nf)
    gc, _ = sem.jacobi_gll(Nc)
    R = sem.interpolation_matrix_1d(gc, gf).T.copy().astype(f32)  # [NqC][NqF]
    qf = r.random(E * (Nf + 1) ** 3).astype(f32)
    pa = r.random(E * (Nf + 1) ** 3).astype(f32)
    return dict(Nf=Nf, Nc=Nc, E=E, R=R, qf=qf, pa=pa)


def linalg_case():
    r = rng(9300)
    N = 4097
    return dict(N=N, w=r.random(N), x=r.random(N), y=r.random(N), Ap=r.random(N), r=r.random(N), alpha=0.37)


def block_case(N, stress, lambda_field):
    """three-field operators: ellipticBlockPartialAxCoeffHex3D (ggeo) / ellipticStressPartialAxCoeffHex3D (vgeo)"""
    E, Np = 4, (N + 1) ** 3
    r = rng(9400 + 10 * N + (5 if stress else 0) + (1 if lambda_field else 0))
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    geo = (r.random((E, 12, Np)) - 0.3) if stress else r.random((E, 7, Np))
    offset, loffset = E * Np + 16, E * Np + 8
    q = r.random(3 * offset)
    if lambda_field:
        lam0, lam1 = r.random(3 * loffset) + 0.5, r.random(3 * loffset)
    else:
        lam0, lam1 = np.zeros(3 * loffset), np.zeros(3 * loffset)
        lam0[[0, loffset, 2 * loffset]] = [1.1, 1.2, 1.3]
        lam1[[0, loffset, 2 * loffset]] = [0.5, 0.6, 0.7]
    el = np.array([2, 0, 3], dtype=np.int32)
    return dict(N=N, E=E, Np=Np, D=D, geo=geo, q=q, el=el, lam0=lam0, lam1=lam1, offset=offset, loffset=loffset,
                stress=stress, lambda_field=lambda_field)


BLOCK_CASES = [(7, False, False), (7, True, False), (3, True, True), (3, False, True)]
AX_CASES = [(7, "d", True), (7, "f", True), (3, "f", True), (1, "f", True), (7, "d", False), (5, "d", True)]
FDM_CASES = [(7, 1), (3, 1), (7, 0)]
TRANSFER_CASES = [(7, 3), (3, 1)]

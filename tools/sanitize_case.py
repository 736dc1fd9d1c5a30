"""Small driver for compute-sanitizer runs (racecheck / synccheck / memcheck): a mesh large enough for the persistent
TMA-ring axhelm to wrap its ring (10 elements per CTA), the gather-scatter kernels, a few PCG iterations and one
multigrid-preconditioned iteration (fusedFDM, transfers, coarse cluster kernel).

    compute-sanitizer --tool racecheck python tools/sanitize_case.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nekrs_b200 import meshgen  # noqa: E402
from nekrs_b200.elliptic import Elliptic, pressure_options  # noqa: E402
from nekrs_b200.lib import DeviceBuffer as DB  # noqa: E402


def main():
    mesh = meshgen.box_mesh(7, (12, 12, 10), kershaw_eps=0.3)  # 1440 elements: ~10 per CTA, ring of 6 stages wraps
    n = mesh.Nelements * mesh.Np
    opts = {"SOLVER": "PCG", "PRECONDITIONER": "NONE", "MAXIMUM ITERATIONS": "3", "SOLVER TOLERANCE": "1e-15"}
    ell = Elliptic(mesh, opts)
    q = np.zeros(ell.fieldOffset)
    q[:n] = np.random.Generator(np.random.PCG64(1)).random(n)
    d_q, d_Aq = DB(like=q), DB.zeros(ell.fieldOffset, np.float64)
    for v in (4, 5, 6):
        ell.set_ax_variant(8, v)
        ell.operator(d_q, d_Aq)
    ell.set_ax_variant(8, -1)
    x = np.zeros(n)
    ell.solve_host(meshgen.kershaw_rhs(mesh), x)
    print("BP5 iterations", ell.Niter)
    small = meshgen.box_mesh(7, (4, 4, 3), kershaw_eps=0.3)
    o2 = pressure_options(**{"MULTIGRID SMOOTHER": "FOURTHOPTCHEBYSHEV+RAS", "MAXIMUM ITERATIONS": "2"})
    e2 = Elliptic(small, o2)
    x2 = np.zeros(small.Nelements * small.Np)
    e2.solve_host(meshgen.kershaw_rhs(small), x2)
    print("BPS5 iterations", e2.Niter)
    # the coarse solve on all SMs (grid barrier) instead of the cluster kernel
    import ctypes
    from nekrs_b200 import lib, ops
    lib.call("nrsb_set_coarse_variant", ctypes.c_int(3))
    x2[:] = 0
    e2.solve_host(meshgen.kershaw_rhs(small), x2)
    lib.call("nrsb_set_coarse_variant", ctypes.c_int(1))
    print("BPS5 iterations (grid coarse kernel)", e2.Niter, "grid", e2.get_int("coarseGridSize"))
    # TMA ring with padded consumer groups (Nq = 6 fp64, Nq = 10 fp32) and the stress operator
    from oracle import sem
    r = np.random.Generator(np.random.PCG64(2))
    for N, dt, E in ((5, np.float64, 1900), (9, np.float32, 700)):
        Np = (N + 1) ** 3
        g, _ = sem.jacobi_gll(N)
        D = sem.dmatrix_1d(g).astype(dt)
        d_Aq = DB.zeros(E * Np, dt)
        ops.ellipticPartialAxCoeffHex3D(N, DB(like=np.arange(E, dtype=np.int32)), DB(like=r.random(E * 7 * Np).astype(dt)),
                                        D, DB(like=r.random(E * Np).astype(dt)), d_Aq, Nelements=E,
                                        lambda0=DB(like=np.ones(1, dt)), variant=4, dtype=dt)
        print("TMA ring N=%d %s" % (N, np.dtype(dt).name), float(np.abs(d_Aq.download(dt)).max()) > 0)
    N, E = 7, 40
    Np = 512
    g, _ = sem.jacobi_gll(N)
    D = sem.dmatrix_1d(g)
    off = E * Np
    lam = np.zeros(3 * off)
    lam[[0, off, 2 * off]] = 1.0
    d_Aq = DB.zeros(3 * off, np.float64)
    ops.ellipticStressPartialAxCoeffHex3D(N, E, off, off, DB(like=np.arange(E, dtype=np.int32)),
                                          DB(like=r.random(E * 12 * Np)), D, DB(like=lam), DB(like=lam),
                                          DB(like=r.random(3 * off)), d_Aq)
    print("stress operator", float(np.abs(d_Aq.download()).max()) > 0)
    # block solver through the handle (per-field masks, three-field gather-scatter, block reductions)
    bm = meshgen.box_mesh(5, (3, 2, 2), kershaw_eps=0.3)
    ob = {"SOLVER": "PCG", "PRECONDITIONER": "JACOBI", "MAXIMUM ITERATIONS": "5", "SOLVER TOLERANCE": "1e-12"}
    eb = Elliptic(bm, ob, poisson=False, Nfields=3, stress_form=True, block_lambda0=[1.0, 1.2, 0.9],
                  block_lambda1=[0.5, 0.6, 0.7], name="velocity")
    offb, nb = eb.fieldOffset, bm.Nelements * bm.Np
    rhs = np.zeros(3 * offb)
    for f in range(3):
        rhs[f * offb:f * offb + nb] = meshgen.kershaw_rhs(bm)
    d_x = DB.zeros(3 * offb, np.float64)
    print("block solver iterations", eb.solve(DB(like=rhs), d_x))


if __name__ == "__main__":
    main()

"""smoke(): one small invocation of the hot path on cuda:0, checked against the oracle."""
import numpy as np


def run():
    from nekrs_b200 import lib, meshgen
    from nekrs_b200.elliptic import Elliptic, pressure_options
    from nekrs_b200.lib import DeviceBuffer as DB
    from oracle import driver
    from oracle.kernels import Orc

    assert lib.device_count() > 0, "smoke() needs a CUDA device (no CPU fallback)"
    lib.call("nrsb_set_device", 0)
    orc = Orc()
    # 1. fused operator Aq = Q Q^T mask A q, N=7 fp64, kershaw mesh
    mesh = meshgen.box_mesh(7, (3, 3, 3), kershaw_eps=0.3)
    n = mesh.Nelements * mesh.Np
    opts = {"SOLVER": "PCG", "PRECONDITIONER": "NONE", "MAXIMUM ITERATIONS": "20", "SOLVER TOLERANCE": "1e-15"}
    ell = Elliptic(mesh, opts)
    ref = driver.OSolver(mesh, opts, orc)
    q = np.random.Generator(np.random.PCG64(1)).random(n)
    out_ref = np.zeros(n)
    ref.ell.operator(q, out_ref)
    qp = np.zeros(ell.fieldOffset)
    qp[:n] = q
    d_Aq = DB.zeros(ell.fieldOffset, np.float64)
    ell.operator(DB(like=qp), d_Aq)
    err = np.max(np.abs(d_Aq.download()[:n] - out_ref)) / np.max(np.abs(out_ref))
    assert err < 1e-12, err
    # 2. BP5: 20 PCG iterations, residual history vs oracle
    rhs = meshgen.kershaw_rhs(mesh)
    ref.solve(rhs, np.zeros(n))
    x = np.zeros(n)
    ell.solve_host(rhs, x)
    h, hr = ell.res_history(), np.array(ref.res_history)
    assert ell.Niter == ref.Niter and np.max(np.abs(h - hr) / hr) < 1e-12
    # 3. BPS5: p-multigrid (RAS + 4th-kind Chebyshev) preconditioned FGMRES, iteration count vs oracle
    opts = pressure_options(**{"MULTIGRID SMOOTHER": "FOURTHOPTCHEBYSHEV+RAS"})
    ell2 = Elliptic(mesh, opts)
    ref2 = driver.OSolver(mesh, opts, orc)
    ref2.solve(rhs, np.zeros(n))
    x = np.zeros(n)
    it = ell2.solve_host(rhs, x)
    assert abs(it - ref2.Niter) <= 1, (it, ref2.Niter)
    print("smoke ok: operator rel err %.2e, BP5 %d its, BPS5 %d its (oracle %d)" % (err, ell.Niter, it, ref2.Niter))


if __name__ == "__main__":
    run()

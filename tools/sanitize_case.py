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


if __name__ == "__main__":
    main()

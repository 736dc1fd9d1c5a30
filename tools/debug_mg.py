"""Developer tool (GPU): walk the V-cycle op by op and compare with the oracle driver."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nekrs_b200 import meshgen  # noqa: E402
from nekrs_b200.elliptic import Elliptic, pressure_options  # noqa: E402
from nekrs_b200.lib import DeviceBuffer as DB  # noqa: E402
from oracle import driver  # noqa: E402
from oracle.kernels import Orc  # noqa: E402


def stat(name, a, b=None):
    s = "%-28s max|a| %.4e nan %d" % (name, np.nanmax(np.abs(a)) if a.size else 0, int(np.isnan(a).sum()))
    if b is not None:
        s += "  | ref max %.4e  relerr %.3e" % (np.max(np.abs(b)), np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
    print(s, flush=True)


def main():
    N, nel, smoother = int(sys.argv[1]), tuple(int(v) for v in sys.argv[2].split(",")), sys.argv[3]
    extra = dict(kv.split("=") for kv in sys.argv[4:])
    mesh = meshgen.box_mesh(N, nel, kershaw_eps=0.3)
    opts = pressure_options(**{"MULTIGRID SMOOTHER": smoother})
    opts.update(extra)
    orc = Orc()
    ref = driver.OSolver(mesh, opts, orc)
    print("oracle ok; levels", [(l.degree, getattr(l, "max_eig_value", None)) for l in ref.levels], flush=True)
    ell = Elliptic(mesh, opts)
    nl = ell.get_int("nLevels")
    print("product levels", [(ell.get_int("level%d:N" % k), ell.get_real("level%d:maxEig" % k)) for k in range(nl)])
    rng = np.random.Generator(np.random.PCG64(2))
    f32 = np.float32
    vec = {}
    for k in range(nl):
        L = ref.levels[k]
        n = L.Nrows
        rhs = rng.random(n).astype(f32)
        rhs[L.ell.mask_ids] = 0
        orc.gs_add(L.ell.ogs, rhs)
        if L.has_smoother:
            x_ref = np.zeros(n, f32)
            L.smooth(rhs.copy(), x_ref, True)
            d_x = DB.zeros(n, f32)
            ell.level_op(k, "smooth", DB(like=rhs), d_x)
            stat("L%d smooth(down)" % k, d_x.download(), x_ref)
            x2 = x_ref.copy()
            L.smooth(rhs.copy(), x2, False)
            d_x2 = DB(like=x_ref)
            ell.level_op(k, "smoothUp", DB(like=rhs), d_x2)
            stat("L%d smooth(up)" % k, d_x2.download(), x2)
        if k > 0:
            Lf = ref.levels[k - 1]
            fine = rng.random(Lf.Nrows).astype(f32)
            c_ref = np.zeros(n, f32)
            L.coarsen(fine.copy(), c_ref, Lf.ell.inv_degree_f, Lf.degree)
            d_c = DB.zeros(n, f32)
            ell.level_op(k, "coarsen", DB(like=fine), d_c)
            stat("L%d coarsen" % k, d_c.download(), c_ref)
            p_ref = fine.copy()
            L.prolongate(c_ref, p_ref, Lf.degree)
            d_p = DB(like=fine)
            ell.level_op(k, "prolongate", DB(like=c_ref), d_p)
            stat("L%d prolongate" % k, d_p.download(), p_ref)
        vec[k] = rhs
    k = nl - 1
    L = ref.levels[k]
    xr = np.zeros(L.Nrows, f32)
    ref.coarse_solve(vec[k].copy(), xr)
    d = DB.zeros(L.Nrows, f32)
    ell.level_op(k, "coarseSolve", DB(like=vec[k]), d)
    stat("coarse solve", d.download(), xr)
    print("coarse its", ell.get_int("coarseIterations"), getattr(ref.coarse, "last_iter", None))
    n = mesh.Nelements * mesh.Np
    r = rng.random(n)
    r[ref.ell.mask_ids] = 0
    orc.gs_add(ref.ell.ogs, r)
    z_ref = np.zeros(n)
    ref.preconditioner(r, z_ref)
    rp = np.zeros(ell.fieldOffset)
    rp[:n] = r
    d_z = DB.zeros(ell.fieldOffset, np.float64)
    ell.preconditioner(DB(like=rp), d_z)
    stat("preconditioner", d_z.download()[:n], z_ref)
    rhs = meshgen.kershaw_rhs(mesh)
    ref.solve(rhs, np.zeros(n))
    x = np.zeros(n)
    try:
        it = ell.solve_host(rhs, x)
    except Exception as e:  # noqa: BLE001
        print("solve failed:", e)
        it = -1
    print("iters", it, ref.Niter)
    print("hist ", ell.res_history()[:8])
    print("ref  ", np.array(ref.res_history[:8]))


if __name__ == "__main__":
    main()

"""kershaw BP5 / BPS5 (examples/kershaw/kershaw.udf:12-135): N=7, n^3 elements per GPU, eps=0.3.

BP5  = PCG, no preconditioner, 1000 iterations max, tol 1e-15 (fixed work) -> (DOF x iter)/s
BPS5 = p-multigrid preconditioned FGMRES, tol 1e-8 relative            -> s/solve, iterations
Timing protocol of the udf: warm-up solve, then the timed solve, min over repetitions; DOF = E*N^3.

    python tools/kershaw_bench.py [--n 20] [--reps 5] [--smoother FOURTHOPTCHEBYSHEV+RAS]
    torchrun ... tools/kershaw_bench.py      (one brick of n^3 elements per rank: weak scaling)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_kershaw(dist, rank, world, local, n=20, N=7, reps=5, bp5_iters=1000, smoother="FOURTHOPTCHEBYSHEV+RAS",
                coarse_tol="1e-1", skip_bp5=False, skip_bps5=False, comm=None):
    """BP5 + BPS5 on one n^3-element kershaw brick per rank; returns the result dict (all ranks).  `dist` is an
    initialised torch.distributed module (or None on one rank), `comm` an existing parallel.Comm to reuse."""
    from nekrs_b200 import lib, meshgen, parallel
    from nekrs_b200.elliptic import Elliptic, pressure_options
    from nekrs_b200.lib import DeviceBuffer as DB
    if comm is None and world > 1:
        comm = parallel.Comm(dist)
    topo_of = (lambda ids: parallel.discover_topology(ids, comm)) if world > 1 else None
    pg = meshgen.brick_partition(world)
    nel = tuple(n * p for p in pg)
    mesh = meshgen.box_mesh(N, nel, kershaw_eps=0.3, rank=rank, nranks=world)
    E, Np = mesh.Nelements, mesh.Np
    dofs = world * E * N ** 3
    rhs = meshgen.kershaw_rhs(mesh)
    res = {"n_gpus": world, "elements_per_gpu": E, "N": N, "dofs": dofs}

    def barrier():
        lib.synchronize()
        if dist is not None:
            dist.barrier()

    def timed_solves(ell, reps_):
        fo = ell.fieldOffset
        rp = np.zeros(fo)
        rp[:E * Np] = rhs
        d_r0 = DB(like=rp)
        d_r, d_x = DB.zeros(fo, np.float64), DB.zeros(fo, np.float64)
        times = []
        for _ in range(reps_):
            for timed in (False, True):  # warm-up solve then timed solve (kershaw.udf:66-85)
                lib.call("nrsb_memcpy_d2d", lib.vp(d_r), lib.vp(d_r0), fo * 8, None)
                lib.call("nrsb_memset", lib.vp(d_x), 0, fo * 8, None)
                barrier()
                prof = timed and os.environ.get("NRSB_PROFILE") == "1"
                if prof:
                    lib.call("nrsb_profiler_start")
                t = time.perf_counter()
                ell.solve(d_r, d_x)
                lib.synchronize()
                dt = time.perf_counter() - t
                if prof:
                    lib.call("nrsb_profiler_stop")
                if timed:
                    times.append(dt)
        if dist is not None:
            import torch
            tt = torch.tensor(times, device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            times = tt.tolist()
        return min(times), ell.Niter

    if not skip_bp5:
        opts = {"SOLVER": "PCG", "PRECONDITIONER": "NONE", "MAXIMUM ITERATIONS": str(bp5_iters),
                "SOLVER TOLERANCE": "1e-15"}
        ell = Elliptic(mesh, opts, comm=comm, topo_of=topo_of)
        ell.autotune()
        t, it = timed_solves(ell, max(2, reps // 2))
        res["bp5"] = {"solve_s": t, "iterations": it, "us_per_iteration": t / it * 1e6,
                      "dof_iter_per_s": dofs * it / t, "dof_iter_per_s_per_gpu": dofs * it / t / world,
                      "GB_s_algorithmic_per_gpu": E * 91936 * it / t / 1e9}
        ell.destroy()
    if not skip_bps5:
        opts = pressure_options(**{"MULTIGRID SMOOTHER": smoother, "COARSE SOLVER TOLERANCE": coarse_tol})
        for kv in os.environ.get("NRSB_EXTRA_OPTS", "").split(";"):  # developer aid: KEY=VALUE;KEY=VALUE
            if "=" in kv:
                opts[kv.split("=", 1)[0].strip().upper()] = kv.split("=", 1)[1].strip()
        ts = time.time()
        ell = Elliptic(mesh, opts, comm=comm, topo_of=topo_of)
        setup_s = time.time() - ts
        t, it = timed_solves(ell, reps)
        res["bps5"] = {"solve_s": t, "iterations": it, "ms_per_iteration": t / max(it, 1) * 1e3,
                       "dof_iter_per_s_per_gpu": dofs * it / t / world,
                       "dof_per_s_per_gpu": dofs / t / world, "setup_s": setup_s, "smoother": smoother,
                       "coarse_iterations_last": ell.get_int("coarseIterations"),
                       "res0": ell.res0Norm, "res": ell.resNorm}
        ell.destroy()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20)
    ap.add_argument("--N", type=int, default=7)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--bp5-iters", type=int, default=1000)
    ap.add_argument("--smoother", default="FOURTHOPTCHEBYSHEV+RAS")
    ap.add_argument("--coarse-tol", default="1e-1")
    ap.add_argument("--skip-bp5", action="store_true")
    ap.add_argument("--skip-bps5", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nekrs_b200 import lib
    lib.call("nrsb_set_device", local)
    res = run_kershaw(dist, rank, world, local, n=args.n, N=args.N, reps=args.reps, bp5_iters=args.bp5_iters,
                      smoother=args.smoother, coarse_tol=args.coarse_tol, skip_bp5=args.skip_bp5,
                      skip_bps5=args.skip_bps5)
    if rank == 0:
        print(json.dumps(res), flush=True)
        if args.out:
            json.dump(res, open(args.out, "w"), indent=1)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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
    from nekrs_b200 import lib, meshgen, parallel
    from nekrs_b200.elliptic import Elliptic, pressure_options
    from nekrs_b200.lib import DeviceBuffer as DB
    lib.call("nrsb_set_device", local)
    comm = parallel.Comm(dist) if world > 1 else None
    topo_of = (lambda ids: parallel.discover_topology(ids, comm)) if world > 1 else None
    pg = meshgen.brick_partition(world)
    nel = tuple(args.n * p for p in pg)
    t0 = time.time()
    mesh = meshgen.box_mesh(args.N, nel, kershaw_eps=0.3, rank=rank, nranks=world)
    E, Np = mesh.Nelements, mesh.Np
    dofs = world * E * args.N ** 3
    rhs = meshgen.kershaw_rhs(mesh)
    res = {"n_gpus": world, "elements_per_gpu": E, "N": args.N, "dofs": dofs}

    def barrier():
        lib.synchronize()
        if dist is not None:
            dist.barrier()

    def timed_solves(ell, reps):
        fo = ell.fieldOffset
        rp = np.zeros(fo)
        rp[:E * Np] = rhs
        d_r0 = DB(like=rp)
        d_r, d_x = DB.zeros(fo, np.float64), DB.zeros(fo, np.float64)
        times = []
        for _ in range(reps):
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

    if not args.skip_bp5:
        opts = {"SOLVER": "PCG", "PRECONDITIONER": "NONE", "MAXIMUM ITERATIONS": str(args.bp5_iters),
                "SOLVER TOLERANCE": "1e-15"}
        ell = Elliptic(mesh, opts, comm=comm, topo_of=topo_of)
        ell.autotune()
        t, it = timed_solves(ell, max(2, args.reps // 2))
        res["bp5"] = {"solve_s": t, "iterations": it, "dof_iter_per_s_per_gpu": dofs * it / t / world,
                      "GB_s_algorithmic_per_gpu": E * 91936 * it / t / 1e9}
        ell.destroy()
    if not args.skip_bps5:
        opts = pressure_options(**{"MULTIGRID SMOOTHER": args.smoother, "COARSE SOLVER TOLERANCE": args.coarse_tol})
        ts = time.time()
        ell = Elliptic(mesh, opts, comm=comm, topo_of=topo_of)
        setup_s = time.time() - ts
        t, it = timed_solves(ell, args.reps)
        res["bps5"] = {"solve_s": t, "iterations": it, "dof_iter_per_s_per_gpu": dofs * it / t / world,
                       "dof_per_s_per_gpu": dofs / t / world, "setup_s": setup_s, "smoother": args.smoother,
                       "coarse_iterations_last": ell.get_int("coarseIterations"),
                       "res0": ell.res0Norm, "res": ell.resNorm}
    if rank == 0:
        print(json.dumps(res), flush=True)
        if args.out:
            json.dump(res, open(args.out, "w"), indent=1)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

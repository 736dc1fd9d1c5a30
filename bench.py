#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native pressure-Poisson path.

Metric (BASELINE.json): GDOF/s of the fused operator  Aq = Q Q^T mask (A q)  (axhelm + on-rank
gather-scatter + halo exchange) at N=7 fp64, DOF = E*N^3 (benchmarkAx.cpp:312, kershaw.udf:99).
Workload at N=1 GPU: BASELINE.json configs[1], "nekrs-bench axhelm + ogs microbench, E=4096 box
mesh, N=7"; for N>1 GPUs every rank owns a 16^3-element brick of one global box (weak scaling)
and the operator includes the NVLink halo exchange.

A "step" is one application of the operator to one E-vector.  One JSON line is printed by rank 0.

  python bench.py --gpus 1 --steps 50 --warmup 5
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference        # the reference's own SERIAL kernels on the host cores
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ORDER = 7
NEL_PER_RANK = (16, 16, 16)  # 4096 elements per GPU
WORKLOAD = ("nekrs-bench axhelm+ogs fused operator, box mesh E=4096 per GPU, N=7, fp64, all-Dirichlet mask "
            "(BASELINE.json configs[1])")


def algorithmic_bytes(N, w):
    """B_Ax = (2+6) Np w ; B_GS = (Nq^3-(Nq-2)^3)(2w+4) ; B_op = sum."""
    Nq = N + 1
    Np = Nq ** 3
    b_ax = 8 * Np * w
    b_gs = (Np - (Nq - 2) ** 3) * (2 * w + 4)
    return b_ax, b_gs


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.proc = None
        self.lines = []

    def run(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = max(smax, float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own SERIAL kernels (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(wid, nel, steps, warmup, reps, seed, barrier, out):
    """One worker = one 'MPI rank' of the SERIAL backend: owns a sub-box, applies
    ellipticPartialAxCoeffHex3D_v0 (reference .c, -O3 -ffast-math build) + CSR gather-scatter + mask.
    All workers start every repetition together (barrier), so they contend for memory bandwidth as MPI ranks
    would; the worker reports the MIN over repetitions of its mean step time."""
    try:
        from nekrs_b200 import meshgen
        from oracle import kernels as K
        from oracle import sem
        N = N_ORDER
        m = meshgen.box_mesh(N, nel)
        E, Np = m.Nelements, m.Np
        orc = K.Orc(fast=True)
        g, _ = sem.jacobi_gll(N)
        D = sem.dmatrix_1d(g)
        r = np.random.Generator(np.random.PCG64(seed))
        ggeo, _ = orc.geometric_factors(E, N, D, sem.jacobi_gll(N)[1], m.x, m.y, m.z)
        ogs_mesh = sem.Ogs(m.global_ids)
        mask_ids, _ = sem.dirichlet_mask_ids(N, E, m.EToB, ogs_mesh, orc)
        ids = m.global_ids.copy()
        ids[mask_ids] = 0
        ogs = sem.Ogs(ids)
        q = r.random(E * Np)
        Aq = np.zeros(E * Np)
        el = np.arange(E, dtype=np.int32)
        if K.ref_available("ax_d_N7_poisson_fast"):
            ax = K.RefAx(N, "d", fast=True)
            kind = "reference"
            fn = lambda: ax(el, ggeo, D, q, Aq)
        else:
            kind = "port"
            fn = lambda: orc.ax(N, el, ggeo, D, q, Aq)

        def step():
            fn()
            orc.mask(mask_ids, Aq)
            orc.gs_add(ogs, Aq)

        barrier.wait()
        for _ in range(warmup):
            step()
        best = float("inf")
        for _ in range(reps):
            barrier.wait()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            best = min(best, (time.perf_counter() - t0) / steps)
        out.put((wid, best, E, kind))
    except Exception as e:  # never leave the others hanging at the barrier
        try:
            barrier.abort()
        except Exception:
            pass
        out.put((wid, float("nan"), 0, "error: %r" % (e,)))


def _worker_grid(cores):
    """Largest py x pz <= cores with py, pz powers of two dividing 16: the 16^3 box is cut into (16, 16/py, 16/pz)
    sub-boxes, one per worker."""
    n = 1
    while n * 2 <= min(cores, 256):
        n *= 2
    py = 1
    while py * py < n:
        py *= 2
    pz = n // py
    return py, pz


def cpu_operator_throughput(steps, warmup, reps=5, cores=None, min_seconds=1.0):
    """GDOF/s of the CPU arm on the full E=4096 workload: one worker process per core (power of two), elements
    split over a 2-D worker grid, every repetition started together, MIN over `reps` repetitions of the MAX
    over workers.  Repetitions are added until at least `min_seconds` have been timed."""
    cores = cores or len(os.sched_getaffinity(0)) or os.cpu_count() or 1
    py, pz = _worker_grid(cores)
    nworkers = py * pz
    ctx = mp.get_context("spawn")
    # a first short pass sizes the repetition count
    def run(steps_, warmup_, reps_):
        barrier = ctx.Barrier(nworkers)
        out = ctx.Queue()
        procs = [ctx.Process(target=_cpu_worker, args=(i, (16, 16 // py, 16 // pz), steps_, warmup_, reps_, 100 + i,
                                                       barrier, out)) for i in range(nworkers)]
        for p in procs:
            p.start()
        res = [out.get(timeout=1800) for _ in procs]
        for p in procs:
            p.join()
        bad = [r for r in res if not (r[1] == r[1])]
        if bad:
            raise RuntimeError("CPU worker failed: %s" % (bad[0][3],))
        return res
    res = run(steps, warmup, reps)
    t = max(r[1] for r in res)
    timed = t * steps * reps
    if timed < min_seconds:
        reps2 = int(np.ceil(min_seconds / max(t * steps, 1e-9)))
        res = run(steps, warmup, max(reps, reps2))
        t = max(r[1] for r in res)
        reps = max(reps, reps2)
    E = sum(r[2] for r in res)
    return E * N_ORDER ** 3 / t / 1e9, nworkers, res[0][3], E, t, reps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    val, cores, kind, E, t, reps = cpu_operator_throughput(steps, warmup)
    sample = ("%d elements (the full E=4096 single-GPU workload; NOT scaled with --gpus) x %d steps x %d repetitions "
              "(min over repetitions of the max over workers), %d worker processes on a %d-core host (no MPI here)"
              % (E, steps, reps, cores, len(os.sched_getaffinity(0))))
    line = {
        "impl": "reference", "metric": "GDOF/s fused Ax+gather-scatter (N=7 fp64)", "value": val, "unit": "GDOF/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements_per_gpu": E, "N": N_ORDER,
                   "arm": "reference SERIAL kernels on the host cores (no GPU work); the CPU work is the 1-GPU "
                          "workload whatever --gpus says"},
        "cpu_baseline": {"value": val, "unit": "GDOF/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def parity_check(bench, rank, world):
    """The operator this run is about to time, checked against the oracle (CPU restatement of the reference's
    serial kernels, pinned to them bit-exact) on the WHOLE mesh: every rank compares its brick.  The oracle is the
    checker here, never the thing measured.  Returns the max relative error of this rank."""
    from nekrs_b200 import meshgen
    from nekrs_b200.lib import DeviceBuffer as DB
    from oracle import driver
    from oracle.kernels import Orc
    N = N_ORDER
    nel = tuple(n * p for n, p in zip(NEL_PER_RANK, bench.proc_grid))
    whole = meshgen.box_mesh(N, nel)
    part = bench.mesh
    Np = part.Np
    if world > 1:
        x0, y0, z0 = part.brick_lo
        ex, ey, ez = part.brick_n
        iz, iy, ix = np.meshgrid(np.arange(z0, z0 + ez), np.arange(y0, y0 + ey), np.arange(x0, x0 + ex), indexing="ij")
        gelem = (ix + nel[0] * (iy + nel[1] * iz)).ravel()
    else:
        gelem = np.arange(part.Nelements)
    gnode = (gelem[:, None] * Np + np.arange(Np)[None, :]).ravel()
    ref = driver.OSolver(whole, {"SOLVER": "PCG", "PRECONDITIONER": "NONE"}, Orc())
    q_glob = np.random.Generator(np.random.PCG64(4242)).random(whole.Nelements * Np)
    out_ref = np.zeros_like(q_glob)
    ref.ell.operator(q_glob, out_ref)
    ell = bench.elliptic
    nloc = part.Nelements * Np
    qp = np.zeros(ell.fieldOffset)
    qp[:nloc] = q_glob[gnode]
    d_q, d_Aq = DB(like=qp), DB.zeros(ell.fieldOffset, np.float64)
    err = 0.0
    for rep in range(3):  # epochs / chunk counters / parity buffers of consecutive launches
        ell.operator(d_q, d_Aq)
        got = d_Aq.download()[:nloc]
        err = max(err, float(np.max(np.abs(got - out_ref[gnode])) / np.max(np.abs(out_ref))))
    return err


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from nekrs_b200 import lib, meshgen
    from nekrs_b200.elliptic import OperatorBench
    lib.call("nrsb_set_device", local_rank)

    N = N_ORDER
    bench = OperatorBench(N, NEL_PER_RANK, rank=rank, nranks=world, dist=dist)
    E, Np = bench.Nelements, bench.Np
    b_ax, b_gs = algorithmic_bytes(N, 8)

    def barrier():
        lib.synchronize()
        if dist is not None:
            dist.barrier()
        lib.synchronize()

    # ---- parity gate before anything is timed: the benchmarked operator (this mesh, this size, this launch
    #      path, incl. the NVLink halo exchange on several ranks) against the oracle, 1e-12 relative (north_star)
    parity = None
    if not args.no_parity:
        parity = parity_check(bench, rank, world)
        if dist is not None:
            import torch
            t = torch.tensor([parity], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            parity = float(t.item())
        if not (parity <= 1e-12):
            raise SystemExit("bench.py: operator parity vs oracle FAILED: relerr %.3e > 1e-12" % parity)

    for _ in range(max(args.warmup, 3)):
        bench.step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)

    # ---- timed region: EXACTLY K operator steps back to back between one event pair on the launching
    #      stream, bracketed by barrier + synchronize; the steps rotate over 3 independent input sets
    #      (453 MB > L2), so no flush kernel sits between them.
    barrier()
    wall0 = time.perf_counter()
    ms_per_step = bench.timed_loop(bench.step, args.steps)
    barrier()
    wall = time.perf_counter() - wall0
    # ---- the dominant kernel alone (axhelm), same rotation, K launches between one event pair
    ms_ax = bench.timed_loop(bench.ax_only, args.steps)
    barrier()
    # ---- single cold launches after an explicit L2 flush (launch + ramp + tail included): reported beside
    cold = [bench.timed_step(flush=True) for _ in range(min(args.steps, 20))]
    ms_step_cold = float(np.median([c[0] for c in cold]))
    ms_ax_cold = float(np.median([c[1] for c in cold]))
    barrier()

    # ---- e2e: host buffers in, host buffers out through the public handle API
    e2e_ms = []
    for _ in range(3):
        bench.e2e_step()
    barrier()
    for _ in range(min(args.steps, 20)):
        e2e_ms.append(bench.e2e_step())
    barrier()
    # the queued host entry: same bytes per step, copies of neighbouring steps overlap (both copy engines)
    bench.e2e_pipelined(4)
    barrier()
    e2e_pipe_ms = bench.e2e_pipelined(min(args.steps, 20))
    barrier()
    clocks = sampler.stop() if sampler else None
    e2e_ms_mean = float(np.mean(e2e_ms))

    # ---- the other half of BASELINE.json's metric: kershaw BP5 (PCG, fixed work) and BPS5 (p-multigrid FGMRES) at
    #      20^3 elements per GPU, kershaw.udf's protocol (warm-up solve, timed solve, min over repetitions)
    kershaw = None
    if not args.no_kershaw:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from kershaw_bench import run_kershaw
        try:
            kershaw = run_kershaw(dist, rank, world, local_rank, n=20, N=N, reps=3, comm=bench.comm)
        except Exception as e:  # the headline line must not be lost to a failure here
            kershaw = {"error": repr(e)}

    if dist is not None:
        import torch
        t = torch.tensor([ms_per_step, ms_ax, e2e_ms_mean, e2e_pipe_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, ms_ax, e2e_ms_mean, e2e_pipe_ms = [float(v) for v in t.tolist()]

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    dofs = world * E * N ** 3
    value = dofs / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    achieved = E * b_ax / (ms_ax * 1e-3) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, kind, Ec, tc, reps = cpu_operator_throughput(20, 3)
        cpu = {"value": v, "unit": "GDOF/s", "cores": cores, "kind": kind,
               "sample": "%d elements (full E=4096 workload) x 20 steps x %d repetitions (min of max over workers), "
                         "%d worker processes, reference SERIAL kernel ellipticPartialAxCoeffHex3D_v0 "
                         "(-O3 -ffast-math) + CSR gather-scatter" % (Ec, reps, cores)}
    line = {
        "metric": "GDOF/s fused Ax+gather-scatter (N=7 fp64)", "value": value, "unit": "GDOF/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements_per_gpu": E, "N": N,
                   "l2": "inputs larger than L2: the K steps rotate over 3 independent input sets of 151 MB each "
                         "(ggeo+q+Aq; 453 MB > 126 MB L2), launched back to back between one CUDA-event pair",
                   "ax_variant": bench.ax_variant, "partition": "brick %s" % (bench.proc_grid,)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": bench.ncu_traffic_bytes(), "kernel": "axhelm (ellipticPartialAxCoeffHex3D)",
                     "algorithmic_bytes_per_launch": E * b_ax, "ms_per_launch": ms_ax, "peak_source": peak_src,
                     "operator_frac": (E * (b_ax + b_gs) / (ms_per_step * 1e-3) / 1e9) / peak,
                     "cold_single_launch": {"ms_ax": ms_ax_cold, "ms_operator": ms_step_cold,
                                            "note": "one launch alone after an explicit L2 flush (write + read "
                                                    "sweep of 256 MiB), event pair around the single launch"}},
        "e2e": {"value": dofs / (e2e_pipe_ms * 1e-3) / 1e9, "unit": "GDOF/s",
                "h2d_bytes_per_step": E * Np * 8, "d2h_bytes_per_step": E * Np * 8, "ms_per_step": e2e_pipe_ms,
                "api": "nrsb_elliptic_operator_host_async x K + nrsb_elliptic_host_wait (pinned host q in, pinned "
                       "host Aq out every step; upload of step k+1 overlaps download of step k-1)",
                "blocking_call": {"value": dofs / (e2e_ms_mean * 1e-3) / 1e9, "ms_per_step": e2e_ms_mean,
                                  "api": "nrsb_elliptic_operator_host (one blocking call per step)"}},
        "parity_relerr": parity, "parity": "fused operator on this mesh vs oracle (whole mesh, every rank its brick), "
                                            "3 consecutive applications, max over ranks; gate 1e-12",
        "kershaw": kershaw,
        "gpu_launches": args.steps * bench.launches_per_step,
        "clocks": clocks, "wall_s": wall,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-kershaw", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

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
def _cpu_worker(args):
    """One worker = one 'MPI rank' of the SERIAL backend: owns a sub-box, applies
    ellipticPartialAxCoeffHex3D_v0 (reference .c, -O3 -ffast-math build) + CSR gather-scatter + mask."""
    nel, steps, warmup, seed = args
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

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps, E, kind


def cpu_operator_throughput(steps, warmup, total_elements=4096, cores=None):
    """GDOF/s of the CPU arm: `cores` workers, elements split evenly, max over workers."""
    cores = cores or os.cpu_count() or 1
    # split the 16^3 box into `cores` slabs along z (power-of-two core counts divide 16)
    nz = 16
    per = max(1, nz // cores) if cores <= nz else 1
    nworkers = min(cores, nz // per)
    jobs = [((16, 16, per), steps, warmup, 100 + i) for i in range(nworkers)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(nworkers) as pool:
        res = pool.map(_cpu_worker, jobs)
    t = max(r[0] for r in res)
    E = sum(r[1] for r in res)
    return E * N_ORDER ** 3 / t / 1e9, nworkers, res[0][2], E, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 3))
    val, cores, kind, E, t = cpu_operator_throughput(steps, warmup)
    sample = "%d elements (full E=4096 workload) x %d steps, %d worker processes (no MPI here)" % (E, steps, cores)
    line = {
        "impl": "reference", "metric": "GDOF/s fused Ax+gather-scatter (N=7 fp64)", "value": val, "unit": "GDOF/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements_per_gpu": E, "N": N_ORDER,
                   "arm": "reference SERIAL kernels on the host cores (no GPU work)"},
        "cpu_baseline": {"value": val, "unit": "GDOF/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
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
        v, cores, kind, Ec, tc = cpu_operator_throughput(10, 1)
        cpu = {"value": v, "unit": "GDOF/s", "cores": cores, "kind": kind,
               "sample": "%d elements (full E=4096 workload) x 10 steps, %d worker processes, reference SERIAL "
                         "kernel ellipticPartialAxCoeffHex3D_v0 (-O3 -ffast-math) + CSR gather-scatter" % (Ec, cores)}
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

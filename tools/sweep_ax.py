"""Developer sweep (GPU): time the axhelm variants and the gather-scatter, like nekrs-bench-axhelm
(src/bench/axHelm/benchmarkAx.cpp:56-426: GDOF/s with DOF = E*N^3, GB/s with (2+6)*Np*w bytes).
Usage: python tools/sweep_ax.py [--orders 7,3] [--elements 4096,16384] [--out file.json]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nekrs_b200 import lib, meshgen, ops  # noqa: E402
from nekrs_b200.lib import DeviceBuffer as DB, Event  # noqa: E402


def time_fn(fn, reps=20, warm=3, flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        if flush:
            lib.l2_flush()
        a, b = Event(), Event()
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_ms(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--orders", default="7")
    ap.add_argument("--elements", default="4096,16384")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    res = []
    for N in [int(v) for v in args.orders.split(",")]:
        Np = (N + 1) ** 3
        g = meshgen.gll_nodes(N)
        # D by barycentric formula (product-side helper lives in the C++ library; here a local copy)
        dx = g[:, None] - g[None, :]
        np.fill_diagonal(dx, 1.0)
        bw = 1.0 / np.prod(dx, axis=1)
        D = (bw[None, :] / bw[:, None]) / dx
        np.fill_diagonal(D, 0.0)
        np.fill_diagonal(D, -D.sum(axis=1))
        for E in [int(v) for v in args.elements.split(",")]:
            for dt in (np.float64, np.float32):
                r = np.random.Generator(np.random.PCG64(1234))
                w = np.dtype(dt).itemsize
                d_g = DB(like=r.random(E * 7 * Np, dtype=np.float32).astype(dt))
                d_q = DB(like=r.random(E * Np, dtype=np.float32).astype(dt))
                d_Aq = DB.zeros(E * Np, dt)
                d_el = DB(like=np.arange(E, dtype=np.int32))
                lam = DB(like=np.ones(1, dtype=dt))
                for variant in (0, 1, 2, 3, 4, 5, 6):
                    fn = lambda: ops.ellipticPartialAxCoeffHex3D(N, d_el, d_g, D, d_q, d_Aq, Nelements=E,
                                                                 lambda0=lam, variant=variant, dtype=dt)
                    med, mn = time_fn(fn)
                    gbs = E * 8 * Np * w / (med * 1e-3) / 1e9
                    gdofs = E * N ** 3 / (med * 1e-3) / 1e9
                    rec = dict(kernel="axhelm", N=N, E=E, dtype=np.dtype(dt).name, variant=variant, ms=med, ms_min=mn,
                               GBs=gbs, GDOFs=gdofs)
                    res.append(rec)
                    print(json.dumps(rec), flush=True)
                # gather-scatter on a cubic box with ~E elements
                n = round(E ** (1 / 3))
                if n ** 3 == E:
                    m = meshgen.box_mesh(N, (n, n, n))
                    o = ops.Ogs(m.global_ids)
                    fn = lambda: o.gather_scatter(d_Aq, dtype=dt)
                    med, mn = time_fn(fn)
                    nshared = 2 * o.nPairs + 4 * o.nQuads + 8 * o.nOcts
                    rec = dict(kernel="gs", N=N, E=E, dtype=np.dtype(dt).name, ms=med, ms_min=mn,
                               GBs_alg=nshared * (2 * w + 4) / (med * 1e-3) / 1e9, rows=[o.nPairs, o.nQuads, o.nOcts, o.nGen])
                    res.append(rec)
                    print(json.dumps(rec), flush=True)
                    # fused operator, no flush between the two kernels
                    def fused():
                        ops.ellipticPartialAxCoeffHex3D(N, d_el, d_g, D, d_q, d_Aq, Nelements=E, lambda0=lam,
                                                        variant=-1, dtype=dt)
                        o.gather_scatter(d_Aq, dtype=dt)
                    med, mn = time_fn(fused)
                    Bop = 8 * Np * w + (Np - (N - 1) ** 3) * (2 * w + 4)
                    rec = dict(kernel="operator", N=N, E=E, dtype=np.dtype(dt).name, ms=med, ms_min=mn,
                               GBs_alg=E * Bop / (med * 1e-3) / 1e9, GDOFs=E * N ** 3 / (med * 1e-3) / 1e9)
                    res.append(rec)
                    print(json.dumps(rec), flush=True)
                    o.destroy()
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

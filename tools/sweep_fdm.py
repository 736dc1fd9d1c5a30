"""Developer sweep (GPU): fusedFDM variants, like nekrs-bench-fdm (src/bench/fdm/benchmarkFDM.cpp:38-296:
bytes = (3 Nqe^3 + 3 Nqe^2) * 4, flops = 12 Nqe^4 + Nqe^3 per element, DOF = (Nqe-1)^3)."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nekrs_b200 import lib, ops  # noqa: E402
from nekrs_b200.lib import DeviceBuffer as DB, Event  # noqa: E402


def time_fn(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
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
    ap.add_argument("--orders", default="7,3")
    ap.add_argument("--elements", default="4096,8000")
    ap.add_argument("--variants", default="0,1")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    f = np.float32
    VARIANTS = [int(v) for v in args.variants.split(",")]
    res = []
    for N in [int(v) for v in args.orders.split(",")]:
        Nq, Nqe = N + 1, N + 3
        for E in [int(v) for v in args.elements.split(",")]:
            r = np.random.Generator(np.random.PCG64(1))
            u = DB(like=r.random(E * Nqe ** 3, dtype=f))
            Sx, Sy, Sz = (DB(like=(r.random(E * Nqe * Nqe, dtype=f) - 0.5)) for _ in range(3))
            invL = DB(like=r.random(E * Nqe ** 3, dtype=f))
            wts = DB(like=r.random(E * Nq ** 3, dtype=f))
            el = DB(like=np.arange(E, dtype=np.int32))
            for restrict in (1, 0):
                Su = DB.zeros(E * (Nq ** 3 if restrict else Nqe ** 3), f)
                for variant in VARIANTS:
                    lib.call("nrsb_set_fdm_variant", ctypes.c_int(variant))
                    fn = lambda: ops.fusedFDM(N, restrict, E, el, Su, Sx, Sy, Sz, invL, wts, u)
                    med, mn = time_fn(fn)
                    nbytes = (3 * Nqe ** 3 + 3 * Nqe ** 2) * 4
                    flops = 12 * Nqe ** 4 + Nqe ** 3
                    rec = dict(kernel="fusedFDM", N=N, E=E, restrict=restrict, variant=variant, ms=med, ms_min=mn,
                               GBs=E * nbytes / (med * 1e-3) / 1e9, TFLOPs=E * flops / (med * 1e-3) / 1e12,
                               GDOFs=E * (Nqe - 1) ** 3 / (med * 1e-3) / 1e9)
                    res.append(rec)
                    print(json.dumps(rec), flush=True)
            w1 = DB.zeros(E * Nqe ** 3, f)
            uu = DB(like=r.random(E * Nq ** 3, dtype=f))
            med, mn = time_fn(lambda: ops.preFDM(N, E, uu, w1))
            rec = dict(kernel="preFDM", N=N, E=E, ms=med, GBs=E * (Nq ** 3 + Nqe ** 3) * 4 / (med * 1e-3) / 1e9)
            res.append(rec)
            print(json.dumps(rec), flush=True)
    lib.call("nrsb_set_fdm_variant", ctypes.c_int(2))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

"""Developer aid: where the fused operator's time goes on one GPU.  Prints us/launch of axhelm alone,
gather-scatter alone (E-vectors L2 resident: 3 x 16.8 MB) and the two back to back, for the current
NRSB_GS_RPT / NRSB_PDL settings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from nekrs_b200 import lib
    from nekrs_b200.elliptic import OperatorBench
    n = int(os.environ.get("NEL", "16"))
    b = OperatorBench(7, (n, n, n))
    for _ in range(10):
        b.step()
    lib.synchronize()
    out = []
    import time
    for name, fn in (("ax", b.ax_only), ("gs", b.gs_only), ("operator", b.step)):
        best = min(b.timed_loop(fn, 60) for _ in range(5))
        lib.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            fn()
        host = (time.perf_counter() - t0) / 200
        lib.synchronize()
        out.append("%s %.2f us (host %.1f us/call)" % (name, best * 1e3, host * 1e6))
    print("RPT=%s PDL=%s E=%d :: %s" % (os.environ.get("NRSB_GS_RPT", "1"), ("gs" if os.environ.get("NRSB_PDL_GS") else "off"),
                                       b.Nelements, " | ".join(out)), flush=True)


if __name__ == "__main__":
    main()

"""Generate tests/golden/ref_kernels.npz: outputs of the REFERENCE's own SERIAL kernels
(kernels/**/*.c of /root/reference, compiled in place into oracle/_ref by oracle/build_ref.py, CI flags -O2)
on the seeded inputs of cases.py.  Run here (where /root/reference exists):

    python tests/golden/make_golden.py

The fixture travels with the repo, so the oracle and the CUDA kernels can be checked against the reference's
output on machines that have neither /root/reference nor oracle/_ref."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import build_ref  # noqa: E402
from oracle import kernels as K  # noqa: E402
from tests.golden import cases  # noqa: E402


def main():
    build_ref.build_ref()
    out = {}
    for N, prec, poisson in cases.AX_CASES:
        c = cases.ax_case(N, prec, poisson)
        a = np.full(c["E"] * c["Np"], -7.0, dtype=c["dt"])
        if poisson:
            K.RefAx(N, prec)(c["el"], c["ggeo"], c["D"], c["q"], a)
        else:
            K.RefAx(N, prec, poisson=False)(c["el"], c["ggeo"], c["D"], c["q"], a, c["lam0"], c["lam1"])
        out["ax_N%d_%s_%s" % (N, prec, "poisson" if poisson else "helmholtz")] = a
    for N, stress, lf in cases.BLOCK_CASES:
        c = cases.block_case(N, stress, lf)
        a = np.full(3 * c["offset"], -7.0)
        ref = (K.RefAxStress if stress else K.RefAxBlock)(N, lf)
        ref(c["el"], c["geo"], c["D"], c["q"], a, c["lam0"], c["lam1"], c["offset"], c["loffset"])
        out["%s_N%d_lambda%d" % ("axstress" if stress else "axblock", N, int(lf))] = a
    for N, restrict in cases.FDM_CASES:
        c = cases.fdm_case(N, restrict)
        ref = K.RefFdm(N, restrict)
        E, Nq, Nqe = c["E"], c["Nq"], c["Nqe"]
        w1 = np.zeros(E * Nqe ** 3, np.float32)
        ref.pre(E, c["u"], w1)
        out["fdm_N%d_r%d_pre" % (N, restrict)] = w1.copy()
        w1 += c["noise"]
        Su = np.zeros(E * (Nq ** 3 if restrict else Nqe ** 3), np.float32)
        ref.fused(E, Su, c["Sx"], c["Sy"], c["Sz"], c["invL"], c["wts"], w1)
        out["fdm_N%d_r%d_fused" % (N, restrict)] = Su.copy()
        if not restrict:
            o = np.zeros(E * Nq ** 3, np.float32)
            ref.post(E, w1, Su, o, c["wts"])
            out["fdm_N%d_r%d_post" % (N, restrict)] = o
    for Nf, Nc in cases.TRANSFER_CASES:
        c = cases.transfer_case(Nf, Nc)
        ref = K.RefTransfer(Nf, Nc)
        a = np.zeros(c["E"] * (Nc + 1) ** 3, np.float32)
        ref.coarsen(c["E"], c["R"], c["qf"], a)
        pa = c["pa"].copy()
        ref.prolongate(c["E"], c["R"], a, pa)
        out["coarsen_%d_%d" % (Nf, Nc)] = a
        out["prolongate_%d_%d" % (Nf, Nc)] = pa
    c = cases.linalg_case()
    ref = K.RefLinAlg("d")
    r = c["r"].copy()
    out["update_pcg_rdotr"] = np.array([ref.update_pcg(c["N"], c["w"], c["Ap"], c["alpha"], r)])
    out["update_pcg_r"] = r
    out["weighted_inner_prod"] = np.array([ref.weighted_inner_prod_many(c["N"], c["w"], c["x"], c["y"])])
    out["weighted_norm2"] = np.array([ref.weighted_norm2_many(c["N"], c["w"], c["x"])])
    y = c["y"].copy()
    ref.axpby_many(c["N"], 0.3, c["x"], -1.7, y)
    out["axpby"] = y
    path = os.path.join(HERE, "ref_kernels.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()

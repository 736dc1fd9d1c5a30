"""The oracle (oracle/nrs_oracle.c) against the committed golden vectors = outputs of the reference's own SERIAL
kernels (tests/golden/make_golden.py).  Bit-exact, CPU only; needs neither /root/reference nor oracle/_ref."""
import os

import numpy as np
import pytest

from tests.golden import cases

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kernels.npz"))


@pytest.mark.parametrize("N,prec,poisson", cases.AX_CASES)
def test_ax(orc, N, prec, poisson):
    c = cases.ax_case(N, prec, poisson)
    b = np.full(c["E"] * c["Np"], -7.0, dtype=c["dt"])
    if poisson:
        orc.ax(N, c["el"], c["ggeo"], c["D"], c["q"], b)
    else:
        orc.ax(N, c["el"], c["ggeo"], c["D"], c["q"], b, c["lam0"], c["lam1"], poisson=False)
    gold = GOLD["ax_N%d_%s_%s" % (N, prec, "poisson" if poisson else "helmholtz")]
    assert np.array_equal(b, gold)
    assert np.all(gold.reshape(c["E"], c["Np"])[1] == -7.0)  # element 1 is not in the list


@pytest.mark.parametrize("N,stress,lf", cases.BLOCK_CASES)
def test_block_and_stress_ax(orc, N, stress, lf):
    c = cases.block_case(N, stress, lf)
    b = np.full(3 * c["offset"], -7.0)
    fn = orc.ax_stress if stress else orc.ax_block
    fn(N, c["el"], c["geo"], c["D"], c["q"], b, c["lam0"], c["lam1"], c["offset"], c["loffset"], lambda_field=lf)
    assert np.array_equal(b, GOLD["%s_N%d_lambda%d" % ("axstress" if stress else "axblock", N, int(lf))])


@pytest.mark.parametrize("N,restrict", cases.FDM_CASES)
def test_fdm(orc, N, restrict):
    c = cases.fdm_case(N, restrict)
    E, Nq, Nqe = c["E"], c["Nq"], c["Nqe"]
    w1 = np.zeros(E * Nqe ** 3, np.float32)
    orc.pre_fdm(E, N, c["u"], w1)
    assert np.array_equal(w1, GOLD["fdm_N%d_r%d_pre" % (N, restrict)])
    w1 += c["noise"]
    Su = np.zeros(E * (Nq ** 3 if restrict else Nqe ** 3), np.float32)
    orc.fused_fdm(E, N, Su, c["Sx"], c["Sy"], c["Sz"], c["invL"], c["wts"], w1, restrict)
    assert np.array_equal(Su, GOLD["fdm_N%d_r%d_fused" % (N, restrict)])
    if not restrict:
        o = np.zeros(E * Nq ** 3, np.float32)
        orc.post_fdm(E, N, w1, Su, o, c["wts"])
        assert np.array_equal(o, GOLD["fdm_N%d_r%d_post" % (N, restrict)])


@pytest.mark.parametrize("Nf,Nc", cases.TRANSFER_CASES)
def test_transfer(orc, Nf, Nc):
    c = cases.transfer_case(Nf, Nc)
    a = np.zeros(c["E"] * (Nc + 1) ** 3, np.float32)
    orc.coarsen(c["E"], Nf, Nc, c["R"], c["qf"], a)
    assert np.array_equal(a, GOLD["coarsen_%d_%d" % (Nf, Nc)])
    pa = c["pa"].copy()
    orc.prolongate(c["E"], Nf, Nc, c["R"], a, pa)
    assert np.array_equal(pa, GOLD["prolongate_%d_%d" % (Nf, Nc)])


def test_linalg(orc):
    c = cases.linalg_case()
    r = c["r"].copy()
    assert orc.update_pcg(c["N"], c["w"], c["Ap"], c["alpha"], r) == GOLD["update_pcg_rdotr"][0]
    assert np.array_equal(r, GOLD["update_pcg_r"])
    assert orc.weighted_inner_prod(c["N"], c["w"], c["x"], c["y"]) == GOLD["weighted_inner_prod"][0]
    assert orc.weighted_norm2_sq(c["N"], c["w"], c["x"]) == GOLD["weighted_norm2"][0]
    y = c["y"].copy()
    orc.axpby(c["N"], 0.3, c["x"], -1.7, y)
    assert np.array_equal(y, GOLD["axpby"])

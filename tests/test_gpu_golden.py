"""The CUDA kernels, through the C ABI, against the committed golden vectors = outputs of the reference's own
SERIAL kernels (tests/golden/make_golden.py).  Tolerances: 1e-12 relative in fp64, 1e-5 in fp32 (BASELINE.json
north_star); differences come from FMA contraction / summation order only."""
import os

import numpy as np
import pytest

from nekrs_b200 import ops
from nekrs_b200.lib import DeviceBuffer as DB
from tests.golden import cases

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kernels.npz"))


def relerr(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("N,prec,poisson", cases.AX_CASES)
@pytest.mark.parametrize("variant", [-1, 0, 2, 5])
def test_ax(N, prec, poisson, variant):
    c = cases.ax_case(N, prec, poisson)
    dt = c["dt"]
    d_Aq = DB(like=np.full(c["E"] * c["Np"], -7.0, dtype=dt))
    ops.ellipticPartialAxCoeffHex3D(N, DB(like=c["el"]), DB(like=c["ggeo"]), c["D"], DB(like=c["q"]), d_Aq,
                                    Nelements=c["el"].size, lambda0=DB(like=c["lam0"]), lambda1=DB(like=c["lam1"]),
                                    poisson=poisson, variant=variant, dtype=dt)
    gold = GOLD["ax_N%d_%s_%s" % (N, prec, "poisson" if poisson else "helmholtz")]
    out = d_Aq.download(dt)
    assert relerr(out, gold) < (1e-12 if prec == "d" else 1e-5)
    assert np.all(out.reshape(c["E"], c["Np"])[1] == -7.0)  # unlisted element untouched


@pytest.mark.parametrize("N,stress,lf", cases.BLOCK_CASES)
def test_block_and_stress_ax(N, stress, lf):
    """The three-field operators against the outputs of the reference's own block / stress kernels."""
    c = cases.block_case(N, stress, lf)
    d_Aq = DB(like=np.full(3 * c["offset"], -7.0))
    fn = ops.ellipticStressPartialAxCoeffHex3D if stress else ops.ellipticBlockPartialAxCoeffHex3D
    fn(N, c["el"].size, c["offset"], c["loffset"], DB(like=c["el"]), DB(like=c["geo"]), c["D"], DB(like=c["lam0"]),
       DB(like=c["lam1"]), DB(like=c["q"]), d_Aq, lambda_field=lf)
    gold = GOLD["%s_N%d_lambda%d" % ("axstress" if stress else "axblock", N, int(lf))]
    out = d_Aq.download()
    assert relerr(out, gold) < 1e-12
    for f in range(3):   # unlisted element untouched, in every field
        assert np.all(out[f * c["offset"]:f * c["offset"] + c["E"] * c["Np"]].reshape(c["E"], c["Np"])[1] == -7.0)


@pytest.mark.parametrize("N,restrict", cases.FDM_CASES)
def test_fdm(N, restrict):
    c = cases.fdm_case(N, restrict)
    E, Nq, Nqe = c["E"], c["Nq"], c["Nqe"]
    f = np.float32
    d_w = DB.zeros(E * Nqe ** 3, f)
    ops.preFDM(N, E, DB(like=c["u"]), d_w)
    pre = GOLD["fdm_N%d_r%d_pre" % (N, restrict)]
    assert np.array_equal(d_w.download(), pre)  # pure data movement: bit-exact
    w_in = (pre + c["noise"]).astype(f)
    nsu = E * (Nq ** 3 if restrict else Nqe ** 3)
    d_Su, d_u = DB.zeros(nsu, f), DB(like=w_in)
    ops.fusedFDM(N, restrict, E, DB(like=np.arange(E, dtype=np.int32)), d_Su, DB(like=c["Sx"]), DB(like=c["Sy"]),
                 DB(like=c["Sz"]), DB(like=c["invL"]), DB(like=c["wts"]), d_u)
    assert relerr(d_Su.download(), GOLD["fdm_N%d_r%d_fused" % (N, restrict)]) < 1e-5


@pytest.mark.parametrize("Nf,Nc", cases.TRANSFER_CASES)
def test_transfer(Nf, Nc):
    c = cases.transfer_case(Nf, Nc)
    f = np.float32
    gold_c = GOLD["coarsen_%d_%d" % (Nf, Nc)]
    d = DB.zeros(gold_c.size, f)
    ops.ellipticPreconCoarsenHex3D(Nf, Nc, c["E"], c["R"], DB(like=c["qf"]), d)
    assert relerr(d.download(), gold_c) < 1e-5
    dp = DB(like=c["pa"])
    ops.ellipticPreconProlongateHex3D(Nf, Nc, c["E"], c["R"], DB(like=gold_c), dp)
    assert relerr(dp.download(), GOLD["prolongate_%d_%d" % (Nf, Nc)]) < 1e-5


def test_linalg():
    c = cases.linalg_case()
    N = c["N"]
    dr = DB(like=c["r"])
    got = ops.ellipticBlockUpdatePCG(N, DB(like=c["w"]), DB(like=c["Ap"]), c["alpha"], dr)
    assert abs(got - GOLD["update_pcg_rdotr"][0]) / GOLD["update_pcg_rdotr"][0] < 1e-12
    assert relerr(dr.download(), GOLD["update_pcg_r"]) < 1e-15
    got = ops.weightedInnerProdMany(N, 1, 0, DB(like=c["w"]), DB(like=c["x"]), DB(like=c["y"]))
    assert abs(got - GOLD["weighted_inner_prod"][0]) / GOLD["weighted_inner_prod"][0] < 1e-12
    got = ops.weightedNorm2Many(N, 1, 0, DB(like=c["w"]), DB(like=c["x"]))
    assert abs(got - GOLD["weighted_norm2"][0]) / GOLD["weighted_norm2"][0] < 1e-12
    dy = DB(like=c["y"])
    ops.axpby(N, 0.3, DB(like=c["x"]), -1.7, dy)
    assert relerr(dy.download(), GOLD["axpby"]) < 1e-15

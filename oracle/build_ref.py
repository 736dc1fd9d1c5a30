"""Build the checker libraries.

TEST INFRASTRUCTURE ONLY (see oracle/nrs_oracle.c header).

1. `oracle/_oracle.so`  <- oracle/nrs_oracle.c  (our restatement; always buildable)
2. `oracle/_ref/*.so`   <- the reference's own SERIAL-backend kernels, compiled
   where they lie under /root/reference/kernels/**.c with the compile-time
   defines OCCA's serial mode would pass (kernel signature
   convention: `device.cpp:167-172`, scalars by const reference).  No reference
   source is copied: a 10-line translation unit `#include`s the files by
   absolute path.  Only possible in the build container; on the GPU box the
   prebuilt `.so` files travel with the snapshot (oracle/_ref is git-ignored,
   not gpurun-ignored).

Flags follow the reference CI (`.github/workflows/ci.yml:19-20`: -O2, no fast-math).
`_fast` twins are built with `-O3 -march=x86-64-v3 -ffast-math` (the reference's production
CPU flags are -O3 -march=native -ffast-math, CMakeLists.txt:107; x86-64-v3 = AVX2+FMA is used
instead of `native` because the library is built in this container and run on the GPU box) used only for the CPU-baseline timing.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
REFDIR = os.path.join(HERE, "_ref")
ORACLE_SO = os.path.join(HERE, "_oracle.so")

# polynomial orders for which reference kernels are prebuilt
AX_ORDERS = (1, 2, 3, 5, 7, 9)
FDM_ORDERS = (1, 3, 7)          # element order N; extended Nq_e = N+3
TRANSFER_PAIRS = ((7, 3), (3, 1), (7, 5), (5, 3), (9, 5), (5, 1))


def _run(cmd):
    subprocess.run(cmd, check=True)


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "nrs_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        _run(["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-o", ORACLE_SO, src, "-lm"])
    fast = ORACLE_SO.replace(".so", "_fast.so")
    if force or not os.path.exists(fast) or os.path.getmtime(fast) < os.path.getmtime(src):
        _run(["gcc", "-O3", "-march=x86-64-v3", "-ffast-math", "-fPIC", "-shared", "-std=c99", "-o", fast, src, "-lm"])
    return ORACLE_SO


def _common_defs(Nq: int):
    return {
        "FUNC(a)": "a", "dlong": "int", "p_blockSize": 256,
        "p_Nq": Nq, "p_Np": Nq ** 3, "p_Nggeo": 7,
        "p_G00ID": 0, "p_G01ID": 1, "p_G11ID": 2, "p_G12ID": 3, "p_G02ID": 4, "p_G22ID": 5, "p_GWJID": 6,
    }


def _compile(name: str, files, defs: dict, fast: bool = False) -> str:
    os.makedirs(REFDIR, exist_ok=True)
    out = os.path.join(REFDIR, name + ("_fast" if fast else "") + ".so")
    if os.path.exists(out):
        return out
    tu = os.path.join(REFDIR, name + ".tu.cpp")
    with open(tu, "w") as f:
        f.write("// generated: includes reference kernels in place, copies nothing\n")
        f.write("#include <cmath>\n#include <cstdlib>\n")
        for fn in files:
            f.write('#include "%s"\n' % os.path.join(REF, "kernels", fn))
    flags = ["-O3", "-march=x86-64-v3", "-ffast-math"] if fast else ["-O2"]
    cmd = ["g++", "-x", "c++", "-std=c++17", "-fPIC", "-shared", "-w"] + flags
    for k, v in defs.items():
        cmd.append("-D%s=%s" % (k, v))
    cmd += ["-o", out, tu]
    _run(cmd)
    os.remove(tu)
    return out


def ref_ax(N: int, prec: str = "d", poisson: bool = True, fast: bool = False) -> str:
    """ellipticPartialAxCoeffHex3D_v0 for order N; prec 'd' (dfloat=double) or 'f'
    (the pfloat instance: same source with dfloat->float, registerEllipticKernels.cpp)."""
    Nq = N + 1
    d = _common_defs(Nq)
    d.update({"dfloat": "double" if prec == "d" else "float", "pfloat": "float", "p_knl": 0,
              "p_lambda": 0, "p_Nfields": 1})
    if poisson:
        d["p_poisson"] = 1
    name = "ax_%s_N%d_%s" % (prec, N, "poisson" if poisson else "helmholtz")
    return _compile(name, ["elliptic/ellipticPartialAxCoeffHex3D.c"], d, fast)


def ref_ax_block(N: int, lambda_field: int) -> str:
    """ellipticBlockPartialAxCoeffHex3D_v0 (three fields, Helmholtz; p_lambda selects per-node coefficients)."""
    Nq = N + 1
    d = _common_defs(Nq)
    d.update({"dfloat": "double", "pfloat": "float", "p_knl": 0, "p_lambda": lambda_field, "p_Nfields": 3})
    return _compile("axblock_d_N%d_lambda%d" % (N, lambda_field), ["elliptic/ellipticBlockPartialAxCoeffHex3D.c"], d)


def ref_ax_stress(N: int, lambda_field: int) -> str:
    """ellipticStressPartialAxCoeffHex3D_v0 (three coupled fields, vgeo ids of src/mesh/mesh3D.h:82-93)."""
    Nq = N + 1
    d = _common_defs(Nq)
    d.update({"dfloat": "double", "pfloat": "float", "p_knl": 0, "p_lambda": lambda_field, "p_Nfields": 3,
              "p_Nvgeo": 12, "p_RXID": 0, "p_RYID": 1, "p_RZID": 2, "p_SXID": 3, "p_SYID": 4, "p_SZID": 5,
              "p_TXID": 6, "p_TYID": 7, "p_TZID": 8, "p_JID": 9, "p_JWID": 10, "p_IJWID": 11})
    return _compile("axstress_d_N%d_lambda%d" % (N, lambda_field), ["elliptic/ellipticStressPartialAxCoeffHex3D.c"], d)


def ref_fdm(N: int, restrict: int, fast: bool = False) -> str:
    Nq = N + 1
    d = _common_defs(Nq)
    d.update({"dfloat": "double", "pfloat": "float", "p_Nq_e": Nq + 2, "p_Np_e": (Nq + 2) ** 3,
              "p_restrict": restrict})
    return _compile("fdm_N%d_r%d" % (N, restrict),
                    ["elliptic/fusedFDM.c", "elliptic/preFDM.c", "elliptic/postFDM.c"], d, fast)


def ref_transfer(Nf: int, Nc: int) -> str:
    d = _common_defs(Nf + 1)
    d.update({"dfloat": "double", "pfloat": "float", "p_NqFine": Nf + 1, "p_NqCoarse": Nc + 1,
              "p_NpFine": (Nf + 1) ** 3, "p_NpCoarse": (Nc + 1) ** 3})
    return _compile("transfer_Nf%d_Nc%d" % (Nf, Nc),
                    ["elliptic/ellipticPreconCoarsenHex3D.c", "elliptic/ellipticPreconProlongateHex3D.c"], d)


def ref_linalg(prec: str = "d", fast: bool = False) -> str:
    """linAlg + Krylov helper kernels; prec 'f' = the p* (pfloat) instances
    (same source with dfloat->float, registerLinAlgKernels.cpp:70-76)."""
    d = _common_defs(8)
    d.update({"dfloat": "double" if prec == "d" else "float", "pfloat": "float", "p_Nfields": 1})
    files = ["linAlg/axpbyMany.c", "linAlg/axpby.c", "linAlg/axmyz.c", "linAlg/axmy.c", "linAlg/innerProd.c",
             "linAlg/weightedInnerProdMany.c", "linAlg/weightedNorm2Many.c", "linAlg/weightedInnerProd.c",
             "linAlg/weightedNorm2.c", "linAlg/norm2.c",
             "elliptic/ellipticBlockUpdatePCG.c", "elliptic/gramSchmidtOrthogonalization.c",
             "elliptic/updatePGMRESSolution.c", "elliptic/fusedResidualAndNorm.c"]
    if prec == "d":
        files += ["core/copyDfloatToPfloat.c", "core/copyPfloatToDfloat.c", "elliptic/axmyzManyPfloat.c",
                  "elliptic/fusedCopyDfloatToPfloat.c"]
    return _compile("linalg_%s" % prec, files, d, fast)


def build_ref(fast: bool = True) -> bool:
    """Compile every prebuilt reference library.  Returns False when /root/reference is absent."""
    if not os.path.isdir(os.path.join(REF, "kernels")):
        return False
    for N in AX_ORDERS:
        ref_ax(N, "d")
        ref_ax(N, "f")
    ref_ax(7, "d", poisson=False)
    for N in (3, 7):
        ref_ax_block(N, 0)
        ref_ax_block(N, 1)
        ref_ax_stress(N, 0)
        ref_ax_stress(N, 1)
    for N in FDM_ORDERS:
        ref_fdm(N, 1)
        ref_fdm(N, 0)
    for nf, nc in TRANSFER_PAIRS:
        ref_transfer(nf, nc)
    ref_linalg("d")
    ref_linalg("f")
    if fast:
        ref_ax(7, "d", fast=True)
        ref_ax(7, "f", fast=True)
        ref_ax(3, "f", fast=True)
        ref_fdm(7, 1, fast=True)
        ref_fdm(3, 1, fast=True)
        ref_linalg("d", fast=True)
    return True


if __name__ == "__main__":
    build_oracle(force="--force" in sys.argv)
    ok = build_ref()
    print("oracle:", ORACLE_SO, "| reference kernels:", "built in " + REFDIR if ok else "SKIPPED (no /root/reference)")

"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/nrsb200.h declares (no compute calls without a GPU)."""
import ctypes
import os

import pytest

from nekrs_b200 import lib


def test_library_exists_and_loads():
    assert os.path.exists(lib.LIB_PATH), "run `python __graft_entry__.py` first"
    L = lib.load()
    assert L.nrsb_version().decode().startswith("nrsb200")


def test_every_declared_symbol_is_exported():
    L = lib.load()
    names = lib.declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/nrsb200.h but not exported: %s" % missing


def test_invalid_arguments_return_codes_not_aborts():
    L = lib.load()
    rc = L.nrsb_ogs_setup(ctypes.c_int32(-1), None, None, None)
    assert rc == -1
    assert b"out is NULL" in L.nrsb_last_error_string() or b"N < 0" in L.nrsb_last_error_string()
    rc = L.nrsb_mask(ctypes.c_int(3), ctypes.c_int32(0), None, None, None)
    assert rc == -1 and b"precision" in L.nrsb_last_error_string()


def test_no_oracle_import_in_product():
    """The product path must not touch oracle/ (parity claims are void otherwise)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for dp, _, fns in os.walk(os.path.join(root, "nekrs_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cpp", ".cuh", ".hpp", ".inc", ".h")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if "oracle" in txt.replace("oracle/ ", ""):
                    for line in txt.splitlines():
                        if "oracle" in line and ("import" in line or "#include" in line or "dlopen" in line
                                                 or "CDLL" in line):
                            bad.append((fn, line.strip()))
    assert not bad, bad


def test_python_mirror_of_the_config_structs_matches_the_library():
    """The ctypes mirrors in nekrs_b200/elliptic.py against sizeof() as compiled into the library (a field added to
    include/nrsb200.h but not to the mirror would shift every later member)."""
    import ctypes as C
    from nekrs_b200 import elliptic as E, lib
    f = lib.load().nrsb_sizeof
    f.restype = C.c_int
    f.argtypes = [C.c_char_p]
    assert f(b"nrsb_elliptic_config") == C.sizeof(E._Config)
    assert f(b"nrsb_shared_topology") == C.sizeof(E._Topo)
    assert f(b"no_such_struct") == 0

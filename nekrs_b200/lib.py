"""ctypes binding of libnrsb200.so (the C ABI of include/nrsb200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is
raised.  The Python layer only moves pointers; all compute is in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnrsb200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "nrsb200.h")

_lib = None


class NrsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nrsb200 error %d: %s" % (code, msg))
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "nekrs_b200: %s not found -- run `python __graft_entry__.py` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.nrsb_last_error_string.restype = C.c_char_p
        _lib.nrsb_version.restype = C.c_char_p
    return _lib


def declared_symbols():
    """Every function name declared in include/nrsb200.h."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nrsb_[A-Za-z0-9_]+)\s*\(", txt)))


def check(rc: int):
    if rc != 0:
        raise NrsbError(rc, load().nrsb_last_error_string().decode())


def call(name, *args):
    check(getattr(load(), name)(*args))


# --------------------------------------------------------------------------- argument helpers
def vp(x):
    """void* from an int address, a DeviceBuffer, a numpy array (host) or None."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, DeviceBuffer):
        return C.c_void_p(x.ptr)
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(C.c_void_p)
    if isinstance(x, C.c_void_p):
        return x
    return C.c_void_p(int(x))


i32, i64, f64, f32 = C.c_int32, C.c_int64, C.c_double, C.c_float


class DeviceBuffer:
    """Owning device allocation (nrsb_malloc / nrsb_free)."""

    def __init__(self, nbytes: int = 0, *, like: np.ndarray | None = None, stream=None):
        self.ptr = 0
        self.nbytes = 0
        self.dtype = None
        self.size = 0
        if like is not None:
            like = np.ascontiguousarray(like)
            nbytes = like.nbytes
            self.dtype = like.dtype
            self.size = like.size
        if nbytes:
            p = C.c_void_p()
            call("nrsb_malloc", C.byref(p), C.c_size_t(nbytes))
            self.ptr = p.value or 0
            self.nbytes = nbytes
        if like is not None and nbytes:
            call("nrsb_memcpy_h2d", vp(self.ptr), vp(like), C.c_size_t(nbytes), vp(stream))
            call("nrsb_stream_synchronize", vp(stream))

    @classmethod
    def empty(cls, n, dtype):
        b = cls(int(n) * np.dtype(dtype).itemsize)
        b.dtype = np.dtype(dtype)
        b.size = int(n)
        return b

    @classmethod
    def zeros(cls, n, dtype):
        b = cls.empty(n, dtype)
        if b.nbytes:
            call("nrsb_memset", vp(b.ptr), C.c_int(0), C.c_size_t(b.nbytes), vp(None))
        return b

    def upload(self, a: np.ndarray, stream=None):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        call("nrsb_memcpy_h2d", vp(self.ptr), vp(a), C.c_size_t(a.nbytes), vp(stream))
        call("nrsb_stream_synchronize", vp(stream))

    def download(self, dtype=None, n=None, stream=None) -> np.ndarray:
        dtype = np.dtype(dtype or self.dtype)
        n = self.nbytes // dtype.itemsize if n is None else n
        out = np.empty(n, dtype=dtype)
        if n:
            call("nrsb_memcpy_d2h", vp(out), vp(self.ptr), C.c_size_t(out.nbytes), vp(stream))
            call("nrsb_stream_synchronize", vp(stream))
        return out

    def free(self):
        if self.ptr:
            load().nrsb_free(vp(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def device_count() -> int:
    n = C.c_int(0)
    try:
        call("nrsb_device_count", C.byref(n))
    except NrsbError:
        return 0
    return n.value


def synchronize():
    call("nrsb_device_synchronize")


class Event:
    def __init__(self):
        self.h = C.c_void_p()
        call("nrsb_event_create", C.byref(self.h))

    def record(self, stream=None):
        call("nrsb_event_record", self.h, vp(stream))

    def synchronize(self):
        call("nrsb_event_synchronize", self.h)

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float(0)
        call("nrsb_event_elapsed_ms", self.h, stop.h, C.byref(ms))
        return ms.value

    def __del__(self):
        try:
            load().nrsb_event_destroy(self.h)
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy array (nrsb_malloc_host)."""

    def __init__(self, n, dtype):
        dtype = np.dtype(dtype)
        self.nbytes = int(n) * dtype.itemsize
        p = C.c_void_p()
        call("nrsb_malloc_host", C.byref(p), C.c_size_t(max(self.nbytes, 1)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_byte * self.nbytes).from_address(self.ptr)).view(dtype)

    def __del__(self):
        try:
            load().nrsb_free_host(vp(self.ptr))
        except Exception:
            pass


def l2_flush(stream=None):
    call("nrsb_l2_flush", vp(stream))

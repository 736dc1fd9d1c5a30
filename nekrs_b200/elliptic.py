"""Host-side mirror of the reference's elliptic interface (elliptic_t / ellipticSolveSetup /
ellipticSolve / ellipticOperator / ellipticPreconditioner, src/solvers/elliptic/elliptic.h:185-238)
over the C ABI.  All numerics run in libnrsb200.so; this module only marshals arguments.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib, meshgen
from .lib import DeviceBuffer, call, vp


class _Topo(C.Structure):
    _fields_ = [("rank", C.c_int), ("nranks", C.c_int), ("nShared", C.c_int64), ("sharedIds", C.c_void_p),
                ("sharerOffsets", C.c_void_p), ("sharerRanks", C.c_void_p)]


class _Config(C.Structure):
    _fields_ = [("N", C.c_int), ("Nelements", C.c_int32), ("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p),
                ("globalIds", C.c_void_p), ("EToB", C.c_void_p), ("topo", C.c_void_p), ("nLevels", C.c_int),
                ("levelOrders", C.c_void_p), ("levelGlobalIds", C.c_void_p), ("levelTopo", C.c_void_p),
                ("options", C.c_char_p), ("poisson", C.c_int), ("lambda0", C.c_double), ("lambda1", C.c_double),
                ("comm", C.c_void_p), ("name", C.c_char_p), ("Nfields", C.c_int), ("stressForm", C.c_int),
                ("blockLambda0", C.c_void_p), ("blockLambda1", C.c_void_p)]


class Topology:
    """Which global ids of this rank also live on other ranks (what gslib's gs_setup discovers in the
    reference, ogsSetup.cpp:150-175).  Built by `parallel.discover_topology`."""

    def __init__(self, rank, nranks, shared_ids, sharer_offsets, sharer_ranks):
        self.shared_ids = np.ascontiguousarray(shared_ids, dtype=np.int64)
        self.sharer_offsets = np.ascontiguousarray(sharer_offsets, dtype=np.int32)
        self.sharer_ranks = np.ascontiguousarray(sharer_ranks, dtype=np.int32)
        self.c = _Topo(rank, nranks, self.shared_ids.size, self.shared_ids.ctypes.data,
                       self.sharer_offsets.ctypes.data, self.sharer_ranks.ctypes.data)


def pressure_options(**kw) -> dict:
    """Option keys the elliptic path reads, with the reference's pressure defaults
    (parReader.cpp:836-841,1099) except the coarse solver: BoomerAMG is outside this path, the
    device Jacobi-PCG stand-in is the default."""
    o = {
        "SOLVER": "PGMRES+FLEXIBLE",
        "PGMRES RESTART": "15",
        "MAXIMUM ITERATIONS": "200",
        "SOLVER TOLERANCE": "1e-8",
        "LINEAR SOLVER STOPPING CRITERION": "RELATIVE",
        "PRECONDITIONER": "MULTIGRID",
        "MULTIGRID SMOOTHER": "FOURTHOPTCHEBYSHEV+ASM",
        "MULTIGRID CHEBYSHEV DEGREE": "3",
        "MULTIGRID CHEBYSHEV MAX EIGENVALUE BOUND FACTOR": "1.1",
        "MULTIGRID COARSE SOLVE": "TRUE",
        "COARSE SOLVER": "JPCG",
        "COARSE SOLVER TOLERANCE": "1e-1",
        "COARSE SOLVER MAXIMUM ITERATIONS": "200",
        "INITIAL GUESS": "PREVIOUS",
    }
    o.update({k.upper(): str(v) for k, v in kw.items()})
    return o


def mg_level_orders(options: dict, N: int):
    """determineMGLevels (MG/determineMGLevels.cpp:58-95)."""
    sched = options.get("MULTIGRID SCHEDULE", "")
    if sched:
        import re
        return sorted({int(v) for v in re.findall(r"p=(\d+)", sched)}, reverse=True)
    schwarz = {1: [1], 2: [2, 1], 3: [3, 1], 4: [4, 2, 1], 5: [5, 3, 1], 6: [6, 3, 1], 7: [7, 3, 1], 8: [8, 5, 1],
               9: [9, 5, 1], 10: [10, 6, 1], 11: [11, 6, 1]}
    other = {1: [1], 2: [2, 1], 3: [3, 1], 4: [4, 2, 1], 5: [5, 3, 1], 6: [6, 4, 2, 1], 7: [7, 5, 3, 1],
             8: [8, 6, 4, 1], 9: [9, 7, 5, 1], 10: [10, 8, 5, 1], 11: [11, 9, 5, 1]}
    sm = options.get("MULTIGRID SMOOTHER", "")
    return (schwarz if ("ASM" in sm or "RAS" in sm) else other)[N]


class Elliptic:
    """elliptic_t handle."""

    def __init__(self, mesh: meshgen.HexMesh, options: dict, *, comm=None, topo_of=None, poisson=True, lambda0=1.0,
                 lambda1=0.0, name="pressure", Nfields=1, stress_form=False, EToB=None, block_lambda0=None,
                 block_lambda1=None):
        """Nfields = 3: block solver (velocity): EToB [Nfields][Nelements][6], one lambda0 / lambda1 per field."""
        self.mesh = mesh
        self.Nfields = Nfields
        self.options = dict(options)
        self.N, self.Np = mesh.N, mesh.Np
        self.Nlocal = mesh.Nelements * mesh.Np
        self._keep = []
        opt_txt = "\n".join("%s=%s" % kv for kv in self.options.items()).encode()
        levels = []
        if "MULTIGRID" in self.options.get("PRECONDITIONER", ""):
            levels = [n for n in mg_level_orders(self.options, mesh.N) if n != mesh.N]
        lvl_ids = [np.ascontiguousarray(meshgen.global_ids_at_order(mesh, n)) for n in levels]
        orders = np.array(levels, dtype=np.int32)
        id_ptrs = (C.c_void_p * max(len(levels), 1))(*[a.ctypes.data for a in lvl_ids])
        topo = topo_of(mesh.global_ids) if topo_of else None
        lvl_topos = [topo_of(a) for a in lvl_ids] if topo_of else []
        topo_ptrs = (C.c_void_p * max(len(levels), 1))(*[C.addressof(t.c) for t in lvl_topos]) if lvl_topos else None
        x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (mesh.x, mesh.y, mesh.z))
        gid = np.ascontiguousarray(mesh.global_ids, dtype=np.int64)
        etob = np.ascontiguousarray(mesh.EToB if EToB is None else EToB, dtype=np.int32).ravel()
        if Nfields > 1 and etob.size == mesh.Nelements * 6:
            etob = np.ascontiguousarray(np.tile(etob, Nfields))
        assert etob.size == max(Nfields, 1) * mesh.Nelements * 6
        bl0 = None if block_lambda0 is None else np.ascontiguousarray(block_lambda0, dtype=np.float64)
        bl1 = None if block_lambda1 is None else np.ascontiguousarray(block_lambda1, dtype=np.float64)
        self._keep += [x, y, z, gid, etob, orders, lvl_ids, id_ptrs, topo, lvl_topos, topo_ptrs, opt_txt, bl0, bl1]
        cfg = _Config(mesh.N, mesh.Nelements, x.ctypes.data, y.ctypes.data, z.ctypes.data, gid.ctypes.data,
                      etob.ctypes.data, C.addressof(topo.c) if topo else None, len(levels),
                      orders.ctypes.data if len(levels) else None,
                      C.cast(id_ptrs, C.c_void_p) if len(levels) else None,
                      C.cast(topo_ptrs, C.c_void_p) if topo_ptrs is not None else None, opt_txt,
                      1 if poisson else 0, lambda0, lambda1, comm.handle if comm is not None else None,
                      name.encode(), Nfields, 1 if stress_form else 0,
                      bl0.ctypes.data if bl0 is not None else None, bl1.ctypes.data if bl1 is not None else None)
        self._h = C.c_void_p()
        call("nrsb_elliptic_setup", C.byref(cfg), C.byref(self._h))
        self.fieldOffset = self.get_int("fieldOffset")
        self.Nmasked = self.get_int("Nmasked")
        self.Niter = 0
        self.res00Norm = self.res0Norm = self.resNorm = 0.0

    # ---- properties
    def get_int(self, key) -> int:
        v = C.c_int64(0)
        call("nrsb_elliptic_get_int", self._h, key.encode(), C.byref(v))
        return v.value

    def get_real(self, key) -> float:
        v = C.c_double(0)
        call("nrsb_elliptic_get_real", self._h, key.encode(), C.byref(v))
        return v.value

    def set_real(self, key, value):
        call("nrsb_elliptic_set_real", self._h, key.encode(), C.c_double(value))

    def get_array(self, key, dtype) -> np.ndarray:
        n = C.c_int64(0)
        call("nrsb_elliptic_get_array", self._h, key.encode(), None, C.c_int64(0), C.byref(n))
        out = np.zeros(n.value, dtype=dtype)
        if n.value:
            call("nrsb_elliptic_get_array", self._h, key.encode(), vp(out), C.c_int64(n.value), C.byref(n))
        return out

    def set_option(self, key, value):
        call("nrsb_elliptic_set_option", self._h, key.encode(), str(value).encode())
        self.options[key.upper()] = str(value).upper()

    def set_coeff_field(self, d_lambda0, d_lambda1=None):
        """ELLIPTIC COEFF FIELD: per-node coefficients (device fp64 buffers owned by the caller)."""
        self._keep += [d_lambda0, d_lambda1]
        call("nrsb_elliptic_set_coeff_field", self._h, vp(d_lambda0), vp(d_lambda1))

    def set_coefficients(self, lambda0, lambda1):
        call("nrsb_elliptic_set_coefficients", self._h, C.c_double(lambda0), C.c_double(lambda1))

    def update_jacobi(self):
        """ellipticUpdateJacobi(elliptic)."""
        call("nrsb_elliptic_update_jacobi", self._h)

    def update_lambda(self):
        """ellipticMultiGridUpdateLambda(elliptic)."""
        call("nrsb_elliptic_update_lambda", self._h)

    def set_ax_variant(self, precision, variant):
        call("nrsb_elliptic_set_ax_variant", self._h, C.c_int(precision), C.c_int(variant))

    def autotune(self):
        a, b = C.c_int(-1), C.c_int(-1)
        call("nrsb_elliptic_autotune", self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    # ---- the reference's entry points
    def solve(self, o_r: DeviceBuffer, o_x: DeviceBuffer):
        """ellipticSolve(elliptic, o_r, o_x)."""
        it = C.c_int(0)
        r00, r0, r = C.c_double(0), C.c_double(0), C.c_double(0)
        call("nrsb_elliptic_solve", self._h, vp(o_r), vp(o_x), C.byref(it), C.byref(r00), C.byref(r0), C.byref(r))
        self.Niter, self.res00Norm, self.res0Norm, self.resNorm = it.value, r00.value, r0.value, r.value
        return self.Niter

    def solve_host(self, rhs: np.ndarray, x: np.ndarray):
        it = C.c_int(0)
        r00, r0, r = C.c_double(0), C.c_double(0), C.c_double(0)
        call("nrsb_elliptic_solve_host", self._h, vp(rhs), vp(x), C.byref(it), C.byref(r00), C.byref(r0),
             C.byref(r))
        self.Niter, self.res00Norm, self.res0Norm, self.resNorm = it.value, r00.value, r0.value, r.value
        return self.Niter

    def operator(self, o_q, o_Aq, *, level=0, precision=8, masked=True):
        """ellipticOperator(elliptic, o_q, o_Aq, precision, masked)."""
        call("nrsb_elliptic_operator", self._h, C.c_int(level), C.c_int(precision), vp(o_q), vp(o_Aq),
             C.c_int(1 if masked else 0))

    def operator_host(self, q: np.ndarray, Aq: np.ndarray):
        call("nrsb_elliptic_operator_host", self._h, vp(q), vp(Aq))

    def operator_host_async(self, q: np.ndarray, Aq: np.ndarray):
        """Queued form: consecutive calls overlap upload / operator / download; finish with host_wait()."""
        call("nrsb_elliptic_operator_host_async", self._h, vp(q), vp(Aq))

    def host_wait(self):
        call("nrsb_elliptic_host_wait", self._h)

    def device_barrier(self):
        """All ranks meet on the handle's stream (device-side all-reduce, no host synchronisation)."""
        call("nrsb_elliptic_device_barrier", self._h)

    def ax(self, o_q, o_Aq, *, level=0, precision=8):
        call("nrsb_elliptic_ax", self._h, C.c_int(level), C.c_int(precision), vp(o_q), vp(o_Aq))

    def gather_scatter(self, o_v, *, level=0, precision=8, masked=True):
        """mask + oogs::startFinish(o_v, ogsAdd): the second half of ellipticOperator alone."""
        call("nrsb_elliptic_gather_scatter", self._h, C.c_int(level), C.c_int(precision), vp(o_v),
             C.c_int(1 if masked else 0))

    def preconditioner(self, o_r, o_z):
        call("nrsb_elliptic_preconditioner", self._h, vp(o_r), vp(o_z))

    def level_op(self, level, op, o_in, o_out):
        call("nrsb_elliptic_level_op", self._h, C.c_int(level), op.encode(), vp(o_in), vp(o_out))

    def res_history(self):
        return self.get_array("resHistory", np.float64)

    def destroy(self):
        if self._h:
            call("nrsb_elliptic_destroy", self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class OperatorBench:
    """The fused operator  Aq = Q Q^T mask (A q)  on a box brick per rank: what bench.py times.

    `nsets` independent copies of the problem (handle with its own geometric factors, q, Aq) are built and
    the timed loops rotate over them: one set is 151 MB at E=4096, N=7 (> the 126 MB L2), three sets are
    453 MB, so every step streams its inputs from HBM without a flush kernel between the steps."""

    def __init__(self, N, nel_per_rank, *, rank=0, nranks=1, dist=None, seed=1234, nsets=3):
        from . import parallel
        self.proc_grid = meshgen.brick_partition(nranks)
        nel = tuple(n * p for n, p in zip(nel_per_rank, self.proc_grid))
        self.mesh = meshgen.box_mesh(N, nel, rank=rank, nranks=nranks)
        self.comm = parallel.Comm(dist) if nranks > 1 else None
        topo_of = (lambda ids: parallel.discover_topology(ids, self.comm)) if nranks > 1 else None
        opts = {"SOLVER": "PCG", "PRECONDITIONER": "NONE", "MAXIMUM ITERATIONS": "100", "SOLVER TOLERANCE": "1e-12"}
        self.Nelements, self.Np = self.mesh.Nelements, self.mesh.Np
        r = np.random.Generator(np.random.PCG64(seed + rank))
        self.sets = []
        for _ in range(nsets):
            ell = Elliptic(self.mesh, opts, comm=self.comm, topo_of=topo_of)
            fo = ell.fieldOffset
            h = np.zeros(fo)
            h[:self.Nelements * self.Np] = r.random(self.Nelements * self.Np)
            self.sets.append((ell, DeviceBuffer(like=h), DeviceBuffer.zeros(fo, np.float64)))
        self.elliptic, self.d_q, self.d_Aq = self.sets[0]
        fo = self.elliptic.fieldOffset
        self.h_q = lib.PinnedBuffer(fo, np.float64)
        self.h_Aq = lib.PinnedBuffer(fo, np.float64)
        self.h_q.array[:] = self.d_q.download()
        self.ax_variant = self.elliptic.autotune()[0]
        for ell, _, _ in self.sets[1:]:
            ell.set_ax_variant(8, self.ax_variant)
        self.launches_per_step = 2  # axhelm (+ in-launch halo push on several ranks), gather-scatter / finish
        self._ev = [lib.Event() for _ in range(3)]
        self._k = 0

    def step(self):
        ell, q, Aq = self.sets[self._k % len(self.sets)]
        self._k += 1
        ell.operator(q, Aq)

    def ax_only(self):
        ell, q, Aq = self.sets[self._k % len(self.sets)]
        self._k += 1
        ell.ax(q, Aq)

    def gs_only(self):
        ell, q, Aq = self.sets[self._k % len(self.sets)]
        self._k += 1
        ell.gather_scatter(Aq)

    def timed_loop(self, fn, steps):
        """ms per call of `fn` over `steps` back-to-back calls (one event pair, launching stream)."""
        e0, e1, _ = self._ev
        # several ranks: the host barrier before this call leaves the ranks tens of microseconds apart, which the first
        # exchange of the timed region would absorb; meet on the device first (no-op on one rank), then start the clock
        self.elliptic.device_barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_ms(e1) / steps

    def timed_step(self, flush=True):
        """(ms for the whole operator, ms for the axhelm launch alone), each timed alone after an L2 flush."""
        if flush:
            lib.l2_flush()
        e0, e1, e2 = self._ev
        e0.record()
        self.elliptic.operator(self.d_q, self.d_Aq)
        e1.record()
        e1.synchronize()
        total = e0.elapsed_ms(e1)
        if flush:
            lib.l2_flush()
        e0.record()
        self.elliptic.ax(self.d_q, self.d_Aq)
        e2.record()
        e2.synchronize()
        return total, e0.elapsed_ms(e2)

    def e2e_step(self):
        e0, e1, _ = self._ev
        e0.record()
        self.elliptic.operator_host(self.h_q.array, self.h_Aq.array)
        e1.record()
        e1.synchronize()
        return e0.elapsed_ms(e1)

    def e2e_pipelined(self, steps):
        """ms per step of `steps` host-to-host operator applications through the queued entry point: every
        step uploads its own q from pinned memory and downloads its own Aq; uploads of step k+1 overlap the
        download of step k-1 (two pinned buffer pairs alternate)."""
        if not hasattr(self, "_hq2"):
            fo = self.elliptic.fieldOffset
            self._hq2 = [self.h_q, lib.PinnedBuffer(fo, np.float64)]
            self._hA2 = [self.h_Aq, lib.PinnedBuffer(fo, np.float64)]
            self._hq2[1].array[:] = self.h_q.array
        e0, e1, _ = self._ev
        e0.record()
        for k in range(steps):
            self.elliptic.operator_host_async(self._hq2[k & 1].array, self._hA2[k & 1].array)
        self.elliptic.host_wait()
        e1.record()
        e1.synchronize()
        return e0.elapsed_ms(e1) / steps

    def ncu_traffic_bytes(self):
        """dram bytes per axhelm launch from the committed ncu capture (profiles/), or None."""
        import json
        import os
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
        try:
            return json.load(open(p)).get("axhelm_fp64_N7_E4096_dram_bytes")
        except Exception:
            return None

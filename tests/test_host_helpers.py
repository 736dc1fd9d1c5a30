"""CPU checks of host-side setup code of the library that needs no GPU (dense eigen-solvers, GLL/D
generation) and of the bootstrap layer (topology discovery over a 2-rank gloo group)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.linalg

from nekrs_b200 import lib
from oracle import sem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dense_helpers():
    rng = np.random.Generator(np.random.PCG64(1))
    for n in (4, 6, 10, 12, 14):
        A = rng.random((n, n))
        A = A + A.T
        B = np.diag(rng.random(n) + 0.5)
        a, b, lam = A.copy(order="F"), B.copy(order="F"), np.zeros(n)
        lib.call("nrsb_sym_generalized_eig", C.c_int(n), lib.vp(a), lib.vp(b), lib.vp(lam))
        w, _ = scipy.linalg.eigh(A, B)
        assert np.max(np.abs(lam - w)) < 1e-11
        Vm = a.reshape(n, n, order="F")
        assert np.max(np.abs(Vm.T @ B @ Vm - np.eye(n))) < 1e-11          # dsygv normalisation
        assert np.max(np.abs(A @ Vm - B @ Vm * lam[None, :])) < 1e-10
        H = np.triu(rng.random((n, n)) - 0.3, -1)
        rho = C.c_double(0)
        Hc = np.asfortranarray(H)
        lib.call("nrsb_spectral_radius", C.c_int(n), lib.vp(Hc), C.byref(rho))
        assert abs(rho.value - np.max(np.abs(np.linalg.eigvals(H)))) < 1e-9
    # not positive definite -> error code, not abort
    a, b, lam = np.eye(3, order="F"), -np.eye(3, order="F"), np.zeros(3)
    with pytest.raises(lib.NrsbError):
        lib.call("nrsb_sym_generalized_eig", C.c_int(3), lib.vp(a), lib.vp(b), lib.vp(lam))


@pytest.mark.parametrize("N", [1, 2, 3, 5, 7, 9, 11])
def test_gll_and_dmatrix_match_oracle(N):
    z, w, D = np.zeros(N + 1), np.zeros(N + 1), np.zeros((N + 1) ** 2)
    lib.call("nrsb_gll", C.c_int(N), lib.vp(z), lib.vp(w), lib.vp(D))
    zr, wr = sem.jacobi_gll(N)
    assert np.max(np.abs(z - zr)) < 1e-15 and np.max(np.abs(w - wr)) < 1e-14
    assert np.max(np.abs(D.reshape(N + 1, N + 1) - sem.dmatrix_1d(zr))) < 1e-12


WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from nekrs_b200 import meshgen, parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
comm = parallel.Comm(dist, create_handle=False)
mesh = meshgen.box_mesh(3, (4, 2, 2), rank=rank, nranks=2)
topo = parallel.discover_topology(mesh.global_ids, comm)
other = meshgen.box_mesh(3, (4, 2, 2), rank=1 - rank, nranks=2)
expect = np.intersect1d(np.unique(mesh.global_ids), np.unique(other.global_ids))
assert np.array_equal(topo.shared_ids, expect), (topo.shared_ids.size, expect.size)
assert expect.size == (2 * 3 + 1) ** 2                      # one shared face plane of 2x2 elements at N=3
assert np.array_equal(topo.sharer_offsets, 2 * np.arange(expect.size + 1))
assert np.array_equal(topo.sharer_ranks.reshape(-1, 2), np.tile([0, 1], (expect.size, 1)))
parts = comm.allgather_array(np.arange(rank + 2, dtype=np.int64))
assert [p.size for p in parts] == [2, 3]
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_topology_discovery_two_ranks_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % {"root": ROOT, "port": 29500 + os.getpid() % 500})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


WORKER_GS = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from nekrs_b200 import meshgen, parallel
from oracle import kernels, sem
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
comm = parallel.Comm(dist, create_handle=False)
orc = kernels.Orc()
N, nel = 3, (4, 3, 2)
whole = meshgen.box_mesh(N, nel, kershaw_eps=0.3)
part = meshgen.box_mesh(N, nel, kershaw_eps=0.3, rank=rank, nranks=2)
topo = parallel.discover_topology(part.global_ids, comm)
# local elements -> global element index (brick partition, lexicographic inside the brick)
x0, y0, z0 = part.brick_lo
ex, ey, ez = part.brick_n
iz, iy, ix = np.meshgrid(np.arange(z0, z0 + ez), np.arange(y0, y0 + ey), np.arange(x0, x0 + ex), indexing="ij")
gelem = (ix + nel[0] * (iy + nel[1] * iz)).ravel()
Np = part.Np
sel = (gelem[:, None] * Np + np.arange(Np)[None, :]).ravel()
assert np.array_equal(whole.global_ids[sel], part.global_ids)     # same numbering on the brick
# reference: gather-scatter of a seeded E-vector on the WHOLE mesh, one rank
v_whole = np.random.Generator(np.random.PCG64(5)).random(whole.global_ids.size)
ref = v_whole.copy()
orc.gs_add(sem.Ogs(whole.global_ids), ref)
# distributed: rows that never leave the rank are summed locally; for shared ids every rank forms its partial
# sum, the partials travel (here: allgather), and are added in ascending rank order (oogs semantics)
v = v_whole[sel].copy()
ids = part.global_ids
shared = np.isin(ids, topo.shared_ids)
loc_ids = ids.copy(); loc_ids[shared] = 0
orc.gs_add(sem.Ogs(loc_ids), v)
part_sum = np.zeros(topo.shared_ids.size)
pos = np.searchsorted(topo.shared_ids, ids[shared])
np.add.at(part_sum, pos, v_whole[sel][shared])
both = comm.allgather_array(part_sum)
ids_all = comm.allgather_array(topo.shared_ids)
assert np.array_equal(ids_all[0], ids_all[1])                     # 2 ranks: both share exactly the same ids
total = both[0] + both[1]
v[shared] = total[pos]
err = np.max(np.abs(v - ref[sel])) / np.max(np.abs(ref))
assert err < 1e-14, err
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_distributed_gather_scatter_two_ranks_gloo(tmp_path):
    """N > 1 host path on CPU: brick partition + topology discovery + partial sums exchanged between two gloo ranks
    reproduce the single-rank oracle gather-scatter on the whole mesh."""
    script = tmp_path / "wgs.py"
    script.write_text(WORKER_GS % {"root": ROOT, "port": 29100 + os.getpid() % 300})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_every_default_mg_schedule_is_instantiated():
    """determineMGLevels (MG/determineMGLevels.cpp:58-95) for N = 1..11, Schwarz and non-Schwarz smoothers: every
    coarsen/prolongate pair and every FDM size the default schedule needs exists in the library, and the C++
    level table equals the harness' (nekrs_b200.elliptic.mg_level_orders)."""
    import ctypes as C
    from nekrs_b200 import lib
    from nekrs_b200.elliptic import mg_level_orders
    L = lib.load()
    for sm in ("FOURTHOPTCHEBYSHEV+ASM", "FOURTHOPTCHEBYSHEV+RAS", "CHEBYSHEV+JACOBI", "DAMPEDJACOBI"):
        opt = ("MULTIGRID SMOOTHER=%s\n" % sm).encode()
        for N in range(1, 12):
            out = (C.c_int * 8)()
            cnt = C.c_int(0)
            assert L.nrsb_mg_levels(C.c_int(N), opt, out, C.c_int(8), C.byref(cnt)) == 0
            assert list(out[:cnt.value]) == mg_level_orders({"MULTIGRID SMOOTHER": sm}, N)
            assert L.nrsb_mg_schedule_supported(C.c_int(N), opt) == 1, (sm, N)
    # user schedule
    opt = b"MULTIGRID SCHEDULE=p=7+degree=3,p=5+degree=3,p=1\nMULTIGRID SMOOTHER=FOURTHOPTCHEBYSHEV+ASM\n"
    out = (C.c_int * 8)()
    cnt = C.c_int(0)
    assert L.nrsb_mg_levels(C.c_int(7), opt, out, C.c_int(8), C.byref(cnt)) == 0
    assert list(out[:cnt.value]) == [7, 5, 1]

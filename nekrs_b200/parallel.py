"""Bootstrap layer for one-process-per-GPU runs: torch.distributed supplies the two host collectives
the C library needs at SETUP time (all-gather of small byte blocks, barrier) and the discovery of
which global ids are shared between ranks.  In the reference this is gslib's gs_setup / crystal
router over MPI (ogsSetup.cpp:150-175, 3rd_party/gslib/src/gs.c).  Nothing here runs on the data
path: halo exchange and scalar all-reduces are device-initiated NVLink stores inside the library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib
from .elliptic import Topology

_ALLGATHER = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_void_p)
_BARRIER = C.CFUNCTYPE(None, C.c_void_p)


class Comm:
    """nrsb_comm_t bound to a torch.distributed process group (any backend)."""

    def __init__(self, dist, group=None, create_handle=True):
        import torch
        self.dist, self.group, self.torch = dist, group, torch
        self.rank = dist.get_rank(group)
        self.nranks = dist.get_world_size(group)
        self._cuda = dist.get_backend(group) == "nccl"

        def allgather(buf, nbytes, _user):
            n = int(nbytes)
            arr = np.ctypeslib.as_array((C.c_uint8 * (n * self.nranks)).from_address(buf))
            self.allgather_into(arr, n)

        def barrier(_user):
            self.barrier()

        self._cb = (_ALLGATHER(allgather), _BARRIER(barrier))
        self.handle = C.c_void_p()
        if create_handle:
            lib.call("nrsb_comm_create", C.c_int(self.rank), C.c_int(self.nranks), self._cb[0], self._cb[1], None,
                     C.byref(self.handle))

    def allgather_into(self, arr: np.ndarray, n: int):
        """arr: uint8[nranks*n]; block `rank` valid on entry, all blocks valid on exit."""
        torch = self.torch
        mine = torch.from_numpy(arr[self.rank * n:(self.rank + 1) * n].copy())
        out = torch.empty(self.nranks * n, dtype=torch.uint8)
        if self._cuda:
            mine, out = mine.cuda(), out.cuda()
        self.dist.all_gather_into_tensor(out, mine, group=self.group) if hasattr(self.dist, "all_gather_into_tensor") \
            and self._cuda else self._allgather_list(out, mine)
        arr[:] = out.cpu().numpy()

    def _allgather_list(self, out, mine):
        parts = [self.torch.empty_like(mine) for _ in range(self.nranks)]
        self.dist.all_gather(parts, mine, group=self.group)
        out.copy_(self.torch.cat(parts))

    def barrier(self):
        if self._cuda:
            self.torch.cuda.synchronize()
        self.dist.barrier(group=self.group)

    def allgather_array(self, a: np.ndarray):
        """variable-length all-gather of a 1-D array -> list of arrays."""
        a = np.ascontiguousarray(a)
        sizes = np.zeros(self.nranks, dtype=np.int64)
        sizes[self.rank] = a.size
        buf = sizes.view(np.uint8)
        self.allgather_into(buf, 8)
        mx = int(sizes.max())
        item = a.dtype.itemsize
        blk = np.zeros(self.nranks * mx * item, dtype=np.uint8)
        blk[self.rank * mx * item:self.rank * mx * item + a.nbytes] = a.view(np.uint8)
        if mx:
            self.allgather_into(blk, mx * item)
        return [blk[r * mx * item:r * mx * item + int(sizes[r]) * item].view(a.dtype).copy() for r in range(self.nranks)]

    def allreduce_sum(self, values: np.ndarray) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        lib.call("nrsb_comm_allreduce_sum", self.handle, C.c_int(v.size), lib.vp(v))
        return v


def discover_topology(ids: np.ndarray, comm: Comm) -> Topology:
    """For every global id of this rank that also appears on another rank: the ascending list of
    ranks holding it.  Hashed rendezvous: id -> rank (id % nranks) collects the holders and answers,
    so no rank ever sees more than its share of the id space (the crystal-router idea of gs_setup)."""
    ids = np.asarray(ids, dtype=np.int64)
    uniq = np.unique(ids[ids > 0])
    P, me = comm.nranks, comm.rank
    dest = uniq % P
    order = np.argsort(dest, kind="stable")
    send = uniq[order]
    counts = np.bincount(dest, minlength=P).astype(np.int64)
    # every rank publishes (counts row, ids sorted by destination)
    all_counts = np.stack(comm.allgather_array(counts))          # [src][dst]
    all_ids = comm.allgather_array(send)
    # ids this rank is the rendezvous point for
    mine_ids, mine_src = [], []
    for src in range(P):
        off = int(all_counts[src, :me].sum())
        n = int(all_counts[src, me])
        mine_ids.append(all_ids[src][off:off + n])
        mine_src.append(np.full(n, src, dtype=np.int32))
    mi = np.concatenate(mine_ids) if mine_ids else np.zeros(0, np.int64)
    ms = np.concatenate(mine_src) if mine_src else np.zeros(0, np.int32)
    o = np.lexsort((ms, mi))
    mi, ms = mi[o], ms[o]
    # keep ids held by >= 2 ranks; answer = (id, holders...) flattened
    if mi.size:
        starts = np.flatnonzero(np.r_[True, mi[1:] != mi[:-1]])
        lens = np.diff(np.r_[starts, mi.size])
        keep = np.repeat(lens > 1, lens)
        ans_id, ans_rank = mi[keep], ms[keep]
    else:
        ans_id, ans_rank = mi, ms
    all_ans_id = comm.allgather_array(ans_id)
    all_ans_rank = comm.allgather_array(ans_rank.astype(np.int32))
    gid = np.concatenate(all_ans_id)
    grk = np.concatenate(all_ans_rank)
    # entries concerning ids I hold
    if gid.size:
        starts = np.flatnonzero(np.r_[True, gid[1:] != gid[:-1]]) if False else None
    sel = np.isin(gid, uniq)
    gid, grk = gid[sel], grk[sel]
    o = np.lexsort((grk, gid))
    gid, grk = gid[o], grk[o]
    if gid.size:
        st = np.flatnonzero(np.r_[True, gid[1:] != gid[:-1]])
        shared = gid[st]
        offsets = np.r_[st, gid.size].astype(np.int32)
    else:
        shared = gid
        offsets = np.zeros(1, dtype=np.int32)
    return Topology(me, P, shared, offsets, grk)

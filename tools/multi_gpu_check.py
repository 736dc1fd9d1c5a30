"""Multi-GPU parity check, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

Every rank owns a brick of one global kershaw box, sets up the solver with the NVLink halo exchange and
the device one-shot all-reduce, and compares against the single-rank ORACLE run on the WHOLE mesh
(restricted to its own elements): fused operator (1e-12), BP5 residual history, BPS5 iteration count.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    # CHECK_SAME_GPU=1: all ranks share GPU 0 (CUDA IPC windows work between processes on one device; the kernels of
    # the ranks are time-sliced, so every cross-rank wait costs a scheduling quantum: correctness only).  NCCL
    # refuses two ranks on one device, the setup collectives then go through gloo.
    same_gpu = os.environ.get("CHECK_SAME_GPU") == "1"
    light = os.environ.get("CHECK_LIGHT") == "1"
    if same_gpu:
        local = 0
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nekrs_b200 import lib, meshgen, parallel
    from nekrs_b200.elliptic import Elliptic, pressure_options
    from nekrs_b200.lib import DeviceBuffer as DB
    from oracle import driver
    from oracle.kernels import Orc
    lib.call("nrsb_set_device", local)
    comm = parallel.Comm(dist)
    ok = True

    def report(name, cond, detail=""):
        nonlocal ok
        ok &= bool(cond)
        print("[rank %d] %-34s %s %s" % (rank, name, "OK" if cond else "FAIL", detail), flush=True)

    # ---- scalar all-reduce through peer windows
    v = comm.allreduce_sum(np.array([rank + 1.0, 10.0 * (rank + 1)]))
    tot = world * (world + 1) / 2
    report("one-shot allreduce", np.allclose(v, [tot, 10 * tot]), str(v))
    for rep in range(50):  # epochs / parity buffers
        v = comm.allreduce_sum(np.array([float(rep + rank)]))
        if abs(v[0] - (world * rep + world * (world - 1) / 2)) > 1e-12:
            report("allreduce epoch %d" % rep, False, str(v))
            break

    N = int(os.environ.get("CHECK_N", "7"))
    nel = tuple(int(x) for x in os.environ.get("CHECK_NEL", "4,4,2").split(","))
    whole = meshgen.box_mesh(N, nel, kershaw_eps=0.3)
    part = meshgen.box_mesh(N, nel, kershaw_eps=0.3, rank=rank, nranks=world)
    Np = part.Np
    # global element index of each local element
    x0, y0, z0 = part.brick_lo
    ex, ey, ez = part.brick_n
    iz, iy, ix = np.meshgrid(np.arange(z0, z0 + ez), np.arange(y0, y0 + ey), np.arange(x0, x0 + ex), indexing="ij")
    gelem = (ix + nel[0] * (iy + nel[1] * iz)).ravel()
    gnode = (gelem[:, None] * Np + np.arange(Np)[None, :]).ravel()
    topo_of = lambda ids: parallel.discover_topology(ids, comm)
    orc = Orc()
    nloc = part.Nelements * Np

    # ---- fused operator + BP5
    opts = {"SOLVER": "PCG", "PRECONDITIONER": "NONE", "MAXIMUM ITERATIONS": "40", "SOLVER TOLERANCE": "1e-15",
            "FUSED HALO AX": "TRUE"}  # the in-kernel halo push (opt-in); the default path is checked below
    ell = Elliptic(part, opts, comm=comm, topo_of=topo_of)
    ref = driver.OSolver(whole, opts, orc)
    report("halo rows present", ell.get_int("NhaloGather") > 0, "NhaloGather=%d overlap=%d" % (
        ell.get_int("NhaloGather"), ell.get_int("overlap")))
    report("mask ids", np.array_equal(np.sort(gnode[ell.get_array("maskIds", np.int32)]),
                                      np.intersect1d(ref.ell.mask_ids, gnode)))
    report("invDegree (global multiplicity)", np.array_equal(ell.get_array("invDegree", np.float64),
                                                              ref.ell.inv_degree[gnode]))
    report("volume", abs(ell.get_real("volume") - ref.mesh.volume) < 1e-12)
    q_glob = np.random.Generator(np.random.PCG64(7)).random(whole.Nelements * Np)
    out_ref = np.zeros_like(q_glob)
    ref.ell.operator(q_glob, out_ref)
    qp = np.zeros(ell.fieldOffset)
    qp[:nloc] = q_glob[gnode]
    d_q, d_Aq = DB(like=qp), DB.zeros(ell.fieldOffset, np.float64)
    for rep in range(3):
        ell.operator(d_q, d_Aq)
    err = np.max(np.abs(d_Aq.download()[:nloc] - out_ref[gnode])) / np.max(np.abs(out_ref))
    report("fused operator (overlap) vs oracle", err < 1e-12, "relerr %.2e" % err)
    # the single-launch path (axhelm with the in-kernel NVLink halo push) against the split path
    # (Ax halo -> mask -> oogs::start -> Ax interior -> oogs::finish): same sums, same order => same bits
    opts_split = dict(opts)
    opts_split["FUSED HALO AX"] = "FALSE"
    opts_split["ENABLE GS COMM OVERLAP"] = "SPLIT"
    ell_split = Elliptic(part, opts_split, comm=comm, topo_of=topo_of)
    d_Aq2 = DB.zeros(ell.fieldOffset, np.float64)
    for rep in range(3):
        ell_split.operator(d_q, d_Aq2)
    same = np.array_equal(d_Aq2.download()[:nloc], d_Aq.download()[:nloc])
    report("in-kernel halo push == split path", same)
    for rep in range(20 if light else 200):  # epochs, counters, parity buffers under back-to-back launches
        ell.operator(d_q, d_Aq)
    report("fused operator after many launches", np.array_equal(d_Aq2.download()[:nloc], d_Aq.download()[:nloc]))
    ell_split.destroy()
    # the unsplit operator: Ax on all elements + the one-launch flag-in-data exchange (oogs_t::exchange_ll)
    opts_ll = dict(opts)
    del opts_ll["FUSED HALO AX"]  # library defaults
    ell_ll = Elliptic(part, opts_ll, comm=comm, topo_of=topo_of)
    d_Aq3 = DB.zeros(ell.fieldOffset, np.float64)
    for rep in range(5):
        ell_ll.operator(d_q, d_Aq3)
    report("one-launch exchange == split path", np.array_equal(d_Aq3.download()[:nloc], d_Aq2.download()[:nloc]))
    qf = q_glob[gnode].astype(np.float32)
    out_ref_f = np.zeros(whole.Nelements * Np, dtype=np.float32)
    ref.ell.operator(q_glob.astype(np.float32), out_ref_f)
    d_qf, d_Aqf = DB(like=np.concatenate([qf, np.zeros(ell.fieldOffset - nloc, np.float32)])), DB.zeros(ell.fieldOffset, np.float32)
    for rep in range(3):
        ell_ll.operator(d_qf, d_Aqf, precision=4)
    errf = np.max(np.abs(d_Aqf.download(np.float32)[:nloc] - out_ref_f[gnode])) / np.max(np.abs(out_ref_f))
    report("fp32 operator (one-launch exchange) vs oracle", errf < 1e-5, "relerr %.2e" % errf)
    ell_ll.destroy()
    # ENABLE GS COMM OVERLAP = TIMED: both forms are measured at setup (10 applications each, max over ranks) and every
    # rank keeps the same one (ellipticSetup.cpp:255-302); whichever it is, the result has the same bits
    opts_t = dict(opts_ll)
    opts_t["ENABLE GS COMM OVERLAP"] = "TIMED"
    ell_t = Elliptic(part, opts_t, comm=comm, topo_of=topo_of)
    tu, ts, pick = ell_t.get_real("overlapTimeUnsplit"), ell_t.get_real("overlapTimeSplit"), ell_t.get_int("splitOverlap")
    picks = comm.allreduce_sum(np.array([float(pick)]))[0]
    report("timed overlap choice", tu > 0 and ts > 0 and pick == (1 if ts < tu else 0) and picks in (0.0, float(world)),
           "unsplit %.1f us split %.1f us -> %s" % (tu * 1e6, ts * 1e6, "split" if pick else "unsplit"))
    d_Aq4 = DB.zeros(ell.fieldOffset, np.float64)
    ell_t.operator(d_q, d_Aq4)
    report("timed choice == split path", np.array_equal(d_Aq4.download()[:nloc], d_Aq2.download()[:nloc]))
    ell_t.destroy()
    rhs_glob = meshgen.kershaw_rhs(whole)
    ref.solve(rhs_glob, np.zeros_like(rhs_glob))
    x = np.zeros(nloc)
    it = ell.solve_host(np.ascontiguousarray(rhs_glob[gnode]), x)
    h, hr = ell.res_history(), np.array(ref.res_history)
    report("BP5 PCG residual history", it == ref.Niter and np.max(np.abs(h - hr) / hr) < 1e-12,
           "its %d/%d maxrel %.2e" % (it, ref.Niter, np.max(np.abs(h - hr) / hr)))

    # ---- block solver (Nfields = 3, stress form, Jacobi-PCG): three-field exchange, per-field masks, block reductions
    if not light:
        lam0b, lam1b = [1.0, 1.3, 0.8], [0.6, 0.5, 0.9]
        opts_b = {"SOLVER": "PCG", "PRECONDITIONER": "JACOBI", "MAXIMUM ITERATIONS": "400", "SOLVER TOLERANCE": "1e-9",
                  "LINEAR SOLVER STOPPING CRITERION": "RELATIVE"}
        ell_b = Elliptic(part, opts_b, comm=comm, topo_of=topo_of, poisson=False, Nfields=3, stress_form=True,
                         block_lambda0=lam0b, block_lambda1=lam1b, name="velocity")
        ref_b = driver.OBlockSolver(whole, opts_b, np.tile(np.asarray(whole.EToB), 3), lam0b, lam1b, orc,
                                    stress_form=True)
        offW, offP = ref_b.fieldOffset, ell_b.fieldOffset
        nW = whole.Nelements * Np
        rb = np.random.Generator(np.random.PCG64(11))
        qW = np.zeros(3 * offW)
        for f in range(3):
            qW[f * offW:f * offW + nW] = rb.random(nW)
        outW = np.zeros(3 * offW)
        ref_b.ell.operator(qW, outW)
        to_part = lambda v: np.concatenate([np.r_[v[f * offW + gnode], np.zeros(offP - nloc)] for f in range(3)])
        d_qb, d_Aqb = DB(like=to_part(qW)), DB.zeros(3 * offP, np.float64)
        for rep in range(3):
            ell_b.operator(d_qb, d_Aqb)
        eb = np.max(np.abs(d_Aqb.download() - to_part(outW))) / np.max(np.abs(outW))
        report("block stress operator vs oracle", eb < 1e-12, "relerr %.2e" % eb)
        rhsW = np.zeros(3 * offW)
        for f in range(3):
            rhsW[f * offW:f * offW + nW] = (f + 1) * rhs_glob
        xW = ref_b.solve(rhsW, np.zeros(3 * offW))
        d_xb = DB.zeros(3 * offP, np.float64)
        itb = ell_b.solve(DB(like=to_part(rhsW)), d_xb)
        exb = np.max(np.abs(d_xb.download() - to_part(xW))) / np.max(np.abs(xW))
        report("block Jacobi-PCG", abs(itb - ref_b.Niter) <= 1 and exb < 1e-7,
               "its %d/%d relerr %.2e" % (itb, ref_b.Niter, exb))
        ell_b.destroy()

    # ---- BPS5: p-multigrid preconditioned FGMRES
    for smoother in (() if light else ("FOURTHOPTCHEBYSHEV+RAS", "FOURTHOPTCHEBYSHEV+ASM")):
        opts = pressure_options(**{"MULTIGRID SMOOTHER": smoother})
        ell2 = Elliptic(part, opts, comm=comm, topo_of=topo_of)
        ref2 = driver.OSolver(whole, opts, orc)
        x_ref = ref2.solve(rhs_glob, np.zeros_like(rhs_glob))
        x = np.zeros(nloc)
        it = ell2.solve_host(np.ascontiguousarray(rhs_glob[gnode]), x)
        e = np.max(np.abs(x - x_ref[gnode])) / np.max(np.abs(x_ref))
        lam = [(ell2.get_real("level%d:maxEig" % k), getattr(ref2.levels[k], "max_eig_value", 0.0)) for k in range(2)]
        report("BPS5 %s" % smoother, abs(it - ref2.Niter) <= 1 and e < 1e-6,
               "its %d/%d relerr %.2e maxEig %s" % (it, ref2.Niter, e, lam))
    dist.barrier()
    flag = torch.tensor([0 if ok else 1], device="cpu" if same_gpu else "cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if flag.item() == 0 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()

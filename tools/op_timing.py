"""Developer aid: pipelined fused-operator timing per kernel on N ranks (NRSB_OP_TIMING=1 prints Ax+push / finish)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nekrs_b200 import lib
    from nekrs_b200.elliptic import OperatorBench
    lib.call("nrsb_set_device", local)
    b = OperatorBench(7, (16, 16, 16), rank=rank, nranks=world, dist=dist)
    for _ in range(10):
        b.step()
    lib.synchronize()
    if dist is not None:
        dist.barrier()
    for rep in range(3):
        ms = b.timed_loop(b.step, 150)
        if rank == 0:
            print("pipelined operator: %.2f us/step" % (ms * 1e3), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

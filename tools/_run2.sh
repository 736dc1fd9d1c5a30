NRSB_OP_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/op_timing.py 2>&1 | grep -E "rank 0|pipelined" | tail -14

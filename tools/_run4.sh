set -x
NG=${NG:-4}
CHECK_NEL=4,4,2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29510 tools/multi_gpu_check.py 2>&1 | grep -E "FAIL|MULTI_GPU_CHECK|Error|error" | head -20
NRSB_OP_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 tools/op_timing.py 2>&1 | grep -E "pipelined" | tail -6
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_n$NG.json; cut -c1-330 gpurun_out/bench_n$NG.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/kershaw_bench.py --reps 3 --skip-bps5 2>&1 | tail -1 | tee gpurun_out/kershaw_n$NG.json

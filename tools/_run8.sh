set -x
NG=${NG:-8}
CHECK_NEL=4,4,4 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29510 tools/multi_gpu_check.py 2>&1 | grep -E "FAIL|MULTI_GPU_CHECK|Error|error" | head -20
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_n$NG.json; cut -c1-330 gpurun_out/bench_n$NG.json
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/kershaw_bench.py --reps 2 --skip-bps5 2>&1 | tail -1 | tee gpurun_out/kershaw_n$NG.json

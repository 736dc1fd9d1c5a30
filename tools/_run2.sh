set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
NG=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_n$NG.json; cut -c1-400 gpurun_out/bench_n$NG.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/kershaw_bench.py --reps 3 2>&1 | tail -1 | tee gpurun_out/kershaw_n$NG.json

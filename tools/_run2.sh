set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -25
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/kershaw_bench.py --reps 3 2>&1 | tail -1 | tee gpurun_out/kershaw_n2.json
timeout 300 python tools/kershaw_bench.py --reps 3 --skip-bp5 2>&1 | tail -1

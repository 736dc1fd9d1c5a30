set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; tail -c 1800 gpurun_out/bench_v2.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ax_tma_kernel -s 12 -c 1 -o gpurun_out/prof_ax_v5 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
NRSB_PROFILE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_bp5.csv python tools/kershaw_bench.py --n 20 --skip-bps5 --bp5-iters 100 --reps 2 > gpurun_out/kb5.log 2>&1
python tools/kershaw_bench.py --n 20 --reps 3 2>&1 | tail -1
ls -la gpurun_out

set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/gs_timing.py 2>&1 | tail -1
NRSB_FUSED_GS=1 python tools/gs_timing.py 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 2500 gpurun_out/bench_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ax_tma_kernel -s 12 -c 1 -o gpurun_out/prof_ax_v5 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:gs_rows_kernel -s 12 -c 1 -o gpurun_out/prof_gs python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out

#!/bin/bash
# Round 2, second session (index sort + per-species sort intervals): measurements on one B200, run under gpurun.
# Everything lands in gpurun_out/; tools/summarize_profiles.py r02b turns the reports into profiles/r02b_*.
set -x
python bench.py --steps 120 --warmup 6 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
python bench.py --workload harris --steps 60 --warmup 6 --e2e 0 --no-cpu-baseline > gpurun_out/r02b_bench_harris.json 2>/dev/null
python bench.py --steps 60 --warmup 6 --e2e 0 --no-cpu-baseline --sort-interval 20 > gpurun_out/r02b_bench_sort20.json 2>/dev/null
VPB_DEFER_SORT=0 python bench.py --steps 60 --warmup 6 --e2e 0 --no-cpu-baseline --sort-interval 20 > gpurun_out/r02b_bench_sort20_nodefer.json 2>/dev/null
python tools/kernel_times.py > gpurun_out/r02b_kernel_times.json 2> gpurun_out/r02b_kernel_times.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r02b_launches.csv \
    python bench.py --steps 14 --warmup 3 --e2e 0 --no-cpu-baseline > gpurun_out/r02b_launch.log 2>&1
# advance_p: the fused sort+push launches of step 6 (electrons, gather variant) and the launches 5 steps after a sort
ncu --set full --clock-control none --import-source on -k regex:advance_p_kernel -s 10 -c 4 -o gpurun_out/r02b_advance_p -f \
    python bench.py --steps 6 --warmup 3 --e2e 0 --no-cpu-baseline > gpurun_out/r02b_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pair_scatter|key_hist" -s 6 -c 3 -o gpurun_out/r02b_sort -f \
    python bench.py --steps 6 --warmup 3 --e2e 0 --no-cpu-baseline > gpurun_out/r02b_ncu_sort.log 2>&1
ls -la gpurun_out/r02b_*.ncu-rep

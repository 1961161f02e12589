#!/bin/bash
# Round-2 measurements on one B200 (run under gpurun); everything lands in gpurun_out/, summaries go to profiles/.
set -x
python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
( time python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python bench.py --workload harris --steps 60 --warmup 5 > gpurun_out/r02_bench_harris.json 2>/dev/null
python tools/kernel_times.py > gpurun_out/r02_kernel_times.json 2> gpurun_out/r02_kernel_times.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 300 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 8 --warmup 3 --e2e 0 --no-cpu-baseline > gpurun_out/r02_launch.log 2>&1
# the dominant kernel at a drifted step (electron and ion launch of step 12 after the sort at step 0)
ncu --set full --clock-control none --import-source on -k regex:advance_p_kernel -s 24 -c 2 -o gpurun_out/r02_advance_p -f \
    python bench.py --steps 12 --warmup 3 --e2e 0 --no-cpu-baseline > gpurun_out/r02_ncu_full.log 2>&1
# the brick/tile kernel (variant 5), same step
VPB_BRICK_DEFAULT=1 VPB_BRICK_CFG=2 ncu --set full --clock-control none --import-source on -k regex:advance_p_brick -s 24 -c 2 \
    -o gpurun_out/r02_advance_p_brick -f python bench.py --steps 12 --warmup 3 --e2e 0 --no-cpu-baseline > gpurun_out/r02_ncu_brick.log 2>&1
ls -la gpurun_out/*.ncu-rep

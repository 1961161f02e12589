#!/bin/bash
# second session (deferred sort, sort intervals 6/12): slab checks and a weak-scaling bench line on N GPUs of one box
N=${1:-2}; STEPS=${2:-48}; CHECKS=${3:-1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$CHECKS" = 1 ]; then
  for extra in "" "--clean" "--steps 25"; do
    echo "== multi_gpu_check $extra"
    timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py $extra 2>&1 | grep -E "multi_gpu_check|FAILED|Error|error" | tail -4
  done
fi
echo "== bench --gpus $N"
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 6 2>/dev/null | tail -1 | tee gpurun_out/r02b_bench_${N}gpu.json | python tools/bench_line.py
if [ "$N" = 8 ]; then
  echo "== Harris-sheet workload, weak"
  timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps $STEPS --warmup 6 --workload harris 2>/dev/null | tail -1 | tee gpurun_out/r02b_bench_harris_${N}gpu.json | python tools/bench_line.py
  echo "== strong scaling: one 128^3 box over $N GPUs"
  timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps $STEPS --warmup 6 --scaling strong 2>/dev/null | tail -1 | tee gpurun_out/r02b_bench_${N}gpu_strong.json | python tools/bench_line.py
fi

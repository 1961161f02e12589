#!/bin/bash
# multi-GPU checks and a weak-scaling bench line on N GPUs of one box: tools/mg_run.sh N [bench steps]
N=${1:-2}; STEPS=${2:-40}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for extra in "" "--clean" "--harris --axis 0" "--steps 25"; do
  echo "== multi_gpu_check $extra"
  timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py $extra 2>&1 | grep -E "multi_gpu_check|FAILED|Error|error" | tail -4
done
echo "== bench --gpus $N (fixed-capacity exchange)"
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/r02_bench_${N}gpu.json | cut -c1-400
echo "== bench --gpus $N (counted exchange, VPB_EXCHANGE_FIXED=0)"
VPB_EXCHANGE_FIXED=0 timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps $STEPS --warmup 5 2>/dev/null | tail -1 | tee gpurun_out/r02_bench_${N}gpu_counted.json | cut -c1-300

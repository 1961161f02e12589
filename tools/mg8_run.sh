#!/bin/bash
# Round-2 multi-GPU records on the GPUs of one box: tools/mg8_run.sh N
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
tools/mg_run.sh $N 40
echo "== strong scaling: one 128^3 box over $N GPUs"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 40 --warmup 5 --scaling strong 2>/dev/null | tail -1 | tee gpurun_out/r02_bench_${N}gpu_strong.json | cut -c1-250
echo "== Harris-sheet workload, weak"
timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 40 --warmup 5 --workload harris 2>/dev/null | tail -1 | tee gpurun_out/r02_bench_harris_${N}gpu.json | cut -c1-250
echo "== C5: 256^3 x 32 ppc x 2 species per GPU (1.07e9 particles per GPU)"
timeout 900 $TR --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 --grid 256 --ppc 32 2>/dev/null | tail -1 | tee gpurun_out/r02_c5_${N}gpu.json | cut -c1-250
echo "== C4: sample/reconnection/reconnection 512x256x256 on $N GPUs through the drop-in seam"
timeout 1500 tools/c4_run.sh $N 512 256 256 16 50 gpu 2> gpurun_out/r02_c4_${N}gpu.err | tee gpurun_out/r02_c4_${N}gpu.json | cut -c1-1200
echo "== C4 on the host cores ($N ranks, CPU reference), 10 steps"
timeout 1500 tools/c4_run.sh $N 512 256 256 16 10 cpu 2> gpurun_out/r02_c4_${N}cpu.err | tee gpurun_out/r02_c4_${N}cpu.json | cut -c1-700

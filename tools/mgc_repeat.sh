#!/bin/bash
# repeat the 2-GPU slab check to measure its flake rate: $1 = repetitions, rest = extra args
n=$1; shift
for i in $(seq 1 $n); do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$i tests/multi_gpu_check.py "$@" 2>&1 | grep -E "multi_gpu_check world|FAILED with|differ" | cut -c1-260
done

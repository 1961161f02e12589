#!/bin/bash
# advance_p ablation around the mover phase (timing only; results are physically invalid with a mask set)
for mask in 0 32 33 1 2; do
  echo "== VPB_DEBUG_SKIP=$mask"
  VPB_DEBUG_SKIP=$mask timeout 300 python bench.py --steps 22 --warmup 3 --e2e 0 --no-cpu-baseline --verbose 2>&1 >/dev/null | grep "advance_p ms" | cut -c1-330
done

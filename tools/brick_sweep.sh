#!/bin/bash
# advance_p brick-kernel configurations, ms per launch through a sort cycle (bench.py --verbose prints every launch)
for cfg in ${CFGS:-0 1 2 4 5 6 8}; do
  echo "== VPB_BRICK_CFG=$cfg"
  VPB_BRICK_CFG=$cfg timeout 300 python bench.py --steps 22 --warmup 3 --e2e 0 --no-cpu-baseline --verbose ${BENCH_ARGS} 2>&1 | grep -E "advance_p ms|roofline" | sed -E 's/.*"roofline": (\{[^}]*\}).*"ms_per_step": ([0-9.]+).*/\1/' | cut -c1-420
done
echo "== linear kernel (variant 2)"
timeout 300 python bench.py --steps 22 --warmup 3 --e2e 0 --no-cpu-baseline --verbose --variant 2 ${BENCH_ARGS} 2>&1 | grep -E "advance_p ms|roofline" | sed -E 's/.*"roofline": (\{[^}]*\}).*/\1/' | cut -c1-420

#!/bin/bash
# advance_p with 3 (default), 2 and 1 resident CTAs per SM, by padding the CTA's shared memory
for kb in 0 56 100; do
  echo "== VPB_EXTRA_SMEM_KB=$kb"
  VPB_EXTRA_SMEM_KB=$kb timeout 300 python bench.py --steps 22 --warmup 3 --e2e 0 --no-cpu-baseline --verbose 2>&1 >/dev/null | grep "advance_p ms" | cut -c1-330
done

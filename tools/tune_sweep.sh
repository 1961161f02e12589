#!/bin/bash
# usage: tools/tune_sweep.sh "<VPB_AP_TUNE values>" [bench args]  — advance_p launch times per tuning value
vals=$1; shift
for t in $vals; do
  echo "== VPB_AP_TUNE=$t"
  VPB_AP_TUNE=$t python bench.py --steps 40 --warmup 3 --e2e 0 --no-cpu-baseline --verbose "$@" 2>&1 | \
    python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('advance_p ms per launch:'):
        x = [float(v) for v in line.split(':')[1].split()]
        e, i = x[0::2], x[1::2]
        print('  electron', ' '.join(f'{v:.2f}' for v in e[:20])); print('  ion     ', ' '.join(f'{v:.2f}' for v in i[:20]))
    elif line.startswith('{'):
        d = json.loads(line); print('  value %.2f G/s  ms/step %.3f  frac %.4f  avg_launch %.3f' % (d['value'] / 1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))
    elif 'rror' in line: print(line.rstrip())
"
done

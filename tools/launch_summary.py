"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv [name-filter ...]
Without filters: one line per kernel (launches, mean, total).  With filters: every matching launch in order."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
filters = sys.argv[2:]
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr):
        continue
    name = r[hdr.index('Kernel Name')].split('(')[0]
    try:
        v = float(r[hdr.index('Metric Value')].replace(',', ''))
    except ValueError:
        continue
    unit = r[hdr.index('Metric Unit')]
    v = v / 1e3 if unit in ('ns', 'nsecond') else v * 1e3 if unit in ('ms', 'msecond') else v
    if filters:
        if any(f in name for f in filters):
            print(f"{name[-60:]:60s} {v:10.1f} us")
    else:
        a = agg.setdefault(name[-70:], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {n:4d} {t / n:10.1f} us avg {t:10.1f} us total")

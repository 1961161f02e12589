"""Turns the ncu reports under gpurun_out/ into the small tracked summaries under profiles/ (run here, no GPU)."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

hdr, units, rows = raw('gpurun_out/r1_advance_p.ncu-rep')
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
summ = []
for r in rows:
    d = {'kernel': r[hdr.index('Kernel Name')], 'note': 'electron launch (20 % of particles cross a face)' if not summ else 'ion launch (4 % cross)'}
    for k in keys:
        if k in hdr:
            d[k] = r[hdr.index(k)] + ' ' + units[hdr.index(k)]
    summ.append(d)
json.dump(summ, open(f'profiles/{tag}_advance_p_ncu_summary.json', 'w'), indent=1)

def gb(s):
    v, u = s.split(); return float(v.replace(',', '')) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[u]
tr = [gb(d['dram__bytes_read.sum']) + gb(d['dram__bytes_write.sum']) for d in summ]
json.dump({'dram_bytes_per_launch': int(sum(tr) / len(tr)),
           'source': f'profiles/{tag}_advance_p_ncu_summary.json (ncu --set full, electron and ion launch of step 4)',
           'particles_per_launch': 134217728}, open('profiles/advance_p_traffic.json', 'w'), indent=1)

rows = list(csv.reader(open('gpurun_out/r1_launches.csv')))
i0 = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[i0]; data = [r for r in rows[i0 + 1:] if len(r) == len(h)]
kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in data:
    name = r[kn].split('(')[0]
    v = float(r[mv].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(r[mu], 1)
    tot[name] = tot.get(name, 0) + v; cnt[name] += 1
T = sum(tot.values())
lines = ["kernel,launches,total_us,share"] + [f"{k},{cnt[k]},{v:.1f},{v / T:.4f}" for k, v in sorted(tot.items(), key=lambda x: -x[1])]
open(f'profiles/{tag}_launch_shares.csv', 'w').write("\n".join(lines) + "\n")
print("\n".join(lines[:8]))
for f, g in (('gpurun_out/r1_launches.csv', f'profiles/{tag}_launches_ncu.csv'), ('gpurun_out/bench_full.json', f'profiles/{tag}_bench.json'),
             ('gpurun_out/kernel_times.json', f'profiles/{tag}_kernel_times.json')):
    if os.path.exists(f):
        open(g, 'w').write(open(f).read())

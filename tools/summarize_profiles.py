"""Turns the ncu reports and bench lines under gpurun_out/ into the small tracked summaries under profiles/ (run here,
no GPU).  Usage: python tools/summarize_profiles.py [tag]   (tag = r02: expects gpurun_out/<tag>_advance_p.ncu-rep,
<tag>_advance_p_brick.ncu-rep, <tag>_launches.csv and the bench JSON lines written by tools/profile_<tag>.sh)"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'smsp__inst_executed_op_global_red.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def summary(rep, notes):
    hdr, units, rows = raw(rep)
    out = []
    for r, note in zip(rows, notes):
        d = {'kernel': r[hdr.index('Kernel Name')], 'note': note}
        for k in KEYS:
            if k in hdr:
                d[k] = (r[hdr.index(k)] + ' ' + units[hdr.index(k)]).strip()
        out.append(d)
    return out


def gb(s):
    v, u = s.split(); return float(v.replace(',', '')) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[u]


NOTES = {
    'r02': ['electron launch, 12 steps after a sort (20 % of particles cross a face per step)', 'ion launch, same step (4 % cross)'],
    # bench.py defaults of the second session: electrons sorted every 6 steps, ions every 12
    'r02b': ['electron launch 5 steps after its sort (the last before the next one; also writes the voxel keys for it)',
             'ion launch, same step (5 steps after its sort)',
             'electron launch of a sort step: the gather variant applies the order of the index sort (loads p[perm[k]], stores position k of the other buffer)',
             'ion launch, 6 steps after its sort'],
}
rep = f'gpurun_out/{tag}_advance_p.ncu-rep'
if os.path.exists(rep):
    summ = summary(rep, NOTES.get(tag, NOTES['r02']))
    json.dump(summ, open(f'profiles/{tag}_advance_p_ncu_summary.json', 'w'), indent=1)
    tr = [gb(d['dram__bytes_read.sum']) + gb(d['dram__bytes_write.sum']) for d in summ]
    if tag == 'r02b':
        # launches of a sort cycle: electrons 5 ordinary + 1 gather of 6, ions 11 ordinary + 1 gather of 12 (the ion gather
        # launch was not captured: the electron launch's excess over an ordinary one is added to an ordinary ion launch)
        e_ord, i_ord, e_gat = tr[0], tr[1], tr[2]
        i_gat = i_ord + (e_gat - e_ord)
        mean = 0.5 * ((5 * e_ord + e_gat) / 6 + (11 * i_ord + i_gat) / 12)
        json.dump({'dram_bytes_per_launch': int(mean),
                   'ordinary_launch': int(0.5 * (e_ord + i_ord)), 'sort_step_launch_electrons': int(e_gat),
                   'source': f'profiles/{tag}_advance_p_ncu_summary.json (ncu --set full): mean over a sort cycle at the bench defaults; '
                             'the launch that applies a sort reads 2.3x the particle bytes (32-byte gathers fetch 64-byte DRAM atoms)',
                   'particles_per_launch': 134217728}, open('profiles/advance_p_traffic.json', 'w'), indent=1)
    else:
        json.dump({'dram_bytes_per_launch': int(sum(tr) / len(tr)),
                   'source': f'profiles/{tag}_advance_p_ncu_summary.json (ncu --set full, mean over the captured electron and ion launches)',
                   'particles_per_launch': 134217728}, open('profiles/advance_p_traffic.json', 'w'), indent=1)
rep = f'gpurun_out/{tag}_sort.ncu-rep'
if os.path.exists(rep):
    json.dump(summary(rep, ['index sort of 134 M particles: voxel keys (left by the previous advance_p) -> digit-0 histogram and per-voxel counts',
                            'first scatter pass: keys -> (voxel, index) pairs by the low 11 bits', 'second pass: pairs -> perm by the high 11 bits']),
              open(f'profiles/{tag}_sort_ncu_summary.json', 'w'), indent=1)
rep = f'gpurun_out/{tag}_advance_p_brick.ncu-rep'
if os.path.exists(rep):
    json.dump(summary(rep, ['brick/tile kernel (variant 5, VPB_BRICK_CFG=2), electron launch, same step as the linear capture', 'ion launch']),
              open(f'profiles/{tag}_advance_p_brick_ncu_summary.json', 'w'), indent=1)

lc = f'gpurun_out/{tag}_launches.csv'
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
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
    open(f'profiles/{tag}_launches_ncu.csv', 'w').write(open(lc).read())
for f in ('bench', 'bench_reference', 'bench_harris', 'bench_sort20', 'bench_sort20_nodefer', 'kernel_times', 'bench_2gpu', 'bench_2gpu_counted', 'bench_4gpu', 'bench_8gpu',
          'bench_8gpu_counted', 'bench_8gpu_strong', 'bench_harris_8gpu', 'c5_1gpu', 'c5_8gpu'):
    src = f'gpurun_out/{tag}_{f}.json'
    if os.path.exists(src) and os.path.getsize(src):
        txt = [ln for ln in open(src).read().splitlines() if ln.startswith('{') or ln.startswith('[')]
        if txt:
            open(f'profiles/{tag}_{f}.json', 'w').write("\n".join(txt) + "\n")

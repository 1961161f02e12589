"""Per-kernel timings of the secondary kernels on the C2 workload (CUDA events, median of repeats), with the
algorithmic-byte roofline of SURVEY.md §8(d).  Usage: python tools/kernel_times.py [--grid 128 --ppc 64]"""
import argparse, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from vpic_b200 import engine as E

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=128); ap.add_argument("--ppc", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
class A: pass
args = A(); args.grid = a.grid; args.ppc = a.ppc; args.uth = 0.18; args.sort_interval = "20"; args.variant = 0; args.scaling = "weak"; args.workload = "uniform"
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
sim = bench.build_sim(args, 0, 1, dev)
for _ in range(3):
    sim.advance()
peak = bench.peaks()[0]
sp = sim.species_list[0]
fa, ia, aa = sim.field_array, sim.interpolator_array, sim.accumulator_array
nv, np_ = sim.g.nv, sp.np

def timeit(fn, reps=a.reps):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

rows = []
def rec(name, ms, bytes_alg):
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    rows.append(dict(kernel=name, ms=round(ms, 4), algorithmic_GB=round(bytes_alg / 1e9, 3), GBps=round(gbs, 1), frac_of_peak=round(gbs / peak, 3)))

# sort_p on drifted (nearly sorted) data, as in the step loop: 68 B per particle algorithmic
for _ in range(5):
    sim.advance()
def index_sort():                      # order only; dropping the pending order leaves the input as it was for the next repeat
    E.sort_p(sp, defer=True); sp._perm_pending = False
rec("sort_p as index sort (order + partition[]; the particles move inside the next advance_p)", timeit(index_sort, reps=3), 68.0 * np_)
rec("sort_p moving the particles (nearly sorted input)", timeit(lambda: E.sort_p(sp), reps=3), 68.0 * np_)
rec("load_interpolator_array", timeit(lambda: E.load_interpolator_array(ia, fa)), (24 + 72) * nv)
rec("unload_accumulator_array", timeit(lambda: E.unload_accumulator_array(fa, aa)), (48 + 24) * nv)
rec("clear_accumulator_array", timeit(lambda: E.clear_accumulator_array(aa)), 48 * nv)
rec("advance_b", timeit(lambda: fa.advance_b(0.5)), 2 * 80 * nv)
rec("vacuum_advance_e (+ghost planes)", timeit(lambda: fa.advance_e(1.0)), 2 * 80 * nv)
rec("clear_jf", timeit(lambda: fa.clear_jf()), 2 * 16 * nv)
rec("synchronize_jf (periodic folds)", timeit(lambda: fa.synchronize_jf()), 0.0 + 6 * 2 * 16 * (a.grid + 1) ** 2)
rec("energy_p", timeit(lambda: E.energy_p(sp, ia)), 32.0 * np_)
rec("vacuum_energy_f", timeit(lambda: fa.energy_f()), 32.0 * nv)
# divergence cleaning, shared-face synchronisation, hydro moments (algorithmic bytes: the field_t slots each one reads / writes)
rec("accumulate_rho_p", timeit(lambda: E.accumulate_rho_p(fa, sp), reps=3), 32.0 * np_)
rec("clear_rhof", timeit(lambda: fa.clear_rhof()), 4 * nv)
rec("synchronize_rho (periodic folds)", timeit(lambda: fa.synchronize_rho()), 6 * 2 * 8 * (a.grid + 1) ** 2)
rec("vacuum_compute_div_e_err (+ghost planes)", timeit(lambda: fa.compute_div_e_err()), (12 + 4 + 4 + 4) * nv)
rec("compute_rms_div_e_err", timeit(lambda: fa.compute_rms_div_e_err()), 4 * nv)
rec("vacuum_clean_div_e", timeit(lambda: fa.clean_div_e()), (4 + 12 + 12) * nv)
rec("compute_div_b_err", timeit(lambda: fa.compute_div_b_err()), (12 + 4) * nv)
rec("compute_rms_div_b_err", timeit(lambda: fa.compute_rms_div_b_err()), 4 * nv)
rec("clean_div_b (+ghost planes)", timeit(lambda: fa.clean_div_b()), (4 + 12 + 12) * nv)
rec("synchronize_tang_e_norm_b", timeit(lambda: fa.synchronize_tang_e_norm_b()), 6 * 2 * 20 * (a.grid + 1) ** 2)
ha = E.HydroArray(sim.g)
rec("accumulate_hydro_p (drifted order)", timeit(lambda: E.accumulate_hydro_p(ha, sp, ia), reps=3), 32.0 * np_ + (72 + 2 * 64) * nv)
E.sort_p(sp)
rec("accumulate_hydro_p (just sorted)", timeit(lambda: E.accumulate_hydro_p(ha, sp, ia), reps=3), 32.0 * np_ + (72 + 2 * 64) * nv)
rec("synchronize_hydro_array (periodic folds)", timeit(lambda: ha.synchronize(fa)), 6 * 2 * 56 * (a.grid + 1) ** 2)
print(json.dumps(dict(workload=f"{a.grid}^3 cells, {a.ppc} ppc, np={np_}", peak_GBps=peak, kernels=rows), indent=1))

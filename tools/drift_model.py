"""Monte-Carlo estimate of accumulator increments per particle through a sort cycle (no GPU needed).

A warp row is 32 consecutive particles of a voxel-sorted array.  Per step a row issues one increment per group of
>= 6 lanes that share a voxel, one per ungrouped in-voxel lane, and two per lane that crosses a face (both streaks go
out as vector REDs).  The model moves thermal particles ballistically (no fields) on a periodic grid and counts those
increments k steps after the sort — the quantity that bounds advance_p in the drifted half of a cycle
(profiles/r01_red_microbench.md: ~60 G increments/s chip-wide).

    python tools/drift_model.py [--uth 0.18 --ppc 64 --steps 20 --grid 12]
"""
import argparse
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--uth", type=float, default=0.18)
ap.add_argument("--ppc", type=int, default=64)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--grid", type=int, default=12)
ap.add_argument("--min-group", type=int, default=6)
a = ap.parse_args()
rng = np.random.default_rng(1)
n = a.grid
npart = n ** 3 * a.ppc
cdt_dx = 0.99 / np.sqrt(3.0)                       # bench.py: dt = 0.99 * Courant, dx = 1, c = 1
pos = rng.random((npart, 3)) * n                   # uniform, like the bench load
u = rng.normal(0, a.uth, (npart, 3))
vel = u / np.sqrt(1 + (u ** 2).sum(1, keepdims=True)) * cdt_dx      # cells per step


def voxel(p):
    c = np.floor(p).astype(np.int64) % n
    return c[:, 0] + n * (c[:, 1] + n * c[:, 2])


order = np.argsort(voxel(pos), kind="stable")      # sort_p
pos, vel = pos[order], vel[order]
rows = npart // 32
print(f"uth={a.uth} ppc={a.ppc} grid={n}^3 min_group={a.min_group}")
print("step  crossing%  groups/row  ungrouped%  increments/particle")
tot = 0.0
for k in range(a.steps):
    v0 = voxel(pos)
    pos = pos + vel
    crossed = voxel(pos) != v0                     # a lane that leaves its voxel this step
    v0r, cr = v0[:rows * 32].reshape(rows, 32), crossed[:rows * 32].reshape(rows, 32)
    inc = groups = ungrouped = 0
    for r in range(0, rows, max(1, rows // 4000)):  # sample of rows
        stay = v0r[r][~cr[r]]
        _, cnt = np.unique(stay, return_counts=True)
        g = int((cnt >= a.min_group).sum())
        s = int(cnt[cnt < a.min_group].sum())
        groups += g; ungrouped += s
        inc += g + s + 2 * int(cr[r].sum())
    nrows = len(range(0, rows, max(1, rows // 4000)))
    ipp = inc / (nrows * 32)
    tot += ipp
    print(f"{k:4d}  {100 * crossed.mean():8.1f}  {groups / nrows:10.2f}  {100 * ungrouped / (nrows * 32):9.1f}  {ipp:10.3f}")
print(f"cycle average: {tot / a.steps:.3f} increments per particle")

"""Multi-GPU parity check (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py [--steps 12]

Every rank runs its slab of a 1 x N x 1 decomposition with particle migration and halo exchange over NCCL, and ALSO
the whole undecomposed problem on its own GPU.  After every step the slab is compared with the matching region of
the single-domain run, order-free as SURVEY.md §8(c) prescribes for different topologies: per-voxel particle counts,
per-voxel sorted particle state, fields and energies within fp32 tolerances (the deposit order differs).
Exit code 0 = pass.  tests/test_gpu_parity.py::test_multi_gpu_slab launches it when >= 2 GPUs are visible.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from vpic_b200 import abi, engine as E, grid as G, parallel, simulation as S   # noqa: E402


def global_particles(rng, n, nx, ny, nz, uth, w):
    import refvpic as R
    return R.random_particles(rng, n, nx, ny, nz, uth=uth, w=w)


def to_local(parts, gny_local, rank, nx, ny, nz):
    """Particles of the global box that live in this rank's y-slab, re-indexed to local voxels."""
    i = parts["i"].astype(np.int64)
    x = i % (nx + 2)
    y = (i // (nx + 2)) % (ny + 2)
    z = i // ((nx + 2) * (ny + 2))
    sel = (y - 1) // gny_local == rank
    out = parts[sel].copy()
    yl = y[sel] - rank * gny_local
    out["i"] = (x[sel] + (nx + 2) * (yl + (gny_local + 2) * z[sel])).astype(np.int32)
    return out


def voxel_table(p, nv, key_of=None):
    """Order-free view: particles sorted by (voxel, then the 7 state words)."""
    order = np.lexsort((p["w"], p["uz"], p["uy"], p["ux"], p["dz"], p["dy"], p["dx"], p["i"]))
    return p[order]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--nx", type=int, default=12)
    ap.add_argument("--ny-per-rank", type=int, default=6)
    ap.add_argument("--nz", type=int, default=8)
    ap.add_argument("--ppc", type=int, default=24)
    args = ap.parse_args()

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    nx, nyl, nz = args.nx, args.ny_per_rank, args.nz
    ny = nyl * world
    dt = G.courant_dt(1, 1, 1, nx, ny, nz, frac=0.97)
    rng = np.random.default_rng(99)
    species = [("electron", -1.0, 1.0, 0.35), ("ion", 1.0, 4.0, 0.12)]
    npart = nx * ny * nz * args.ppc
    loads = [global_particles(rng, npart, nx, ny, nz, uth, 1.0 / args.ppc) for _, _, _, uth in species]
    seed_fields = np.zeros(((nx + 2) * (ny + 2) * (nz + 2), 20), np.float32)   # fields start at zero

    # --- single-domain run (every rank has its own copy) ---
    gg = G.partition_periodic_box(0, 0, 0, nx, ny, nz, nx, ny, nz, 1, 1, 1, dt=dt)
    dgg = E.DeviceGrid(gg, dev)
    ref = S.Simulation(dgg)
    for (name, q, m, _), load in zip(species, loads):
        sp = ref.define_species(name, q, m, npart, npart, sort_interval=5)
        sp.set_particles(load)
    ref.initialize()

    # --- slab run ---
    gl = G.partition_periodic_box(0, 0, 0, nx, ny, nz, nx, ny, nz, 1, world, 1, rank=rank, dt=dt)
    dgl = E.DeviceGrid(gl, dev)
    sim = S.Simulation(dgl, exchange=parallel.SlabExchange(dgl, axis=1))
    for (name, q, m, _), load in zip(species, loads):
        mine = to_local(load, nyl, rank, nx, ny, nz)
        sp = sim.define_species(name, q, m, int(npart / world * 1.6) + 64, npart, sort_interval=5)
        sp.set_particles(mine)
    sim.initialize()

    worst = dict(count=0, part=0.0, field=0.0, energy=0.0)
    ok = True
    for step in range(args.steps):
        ref.advance()
        sim.advance()
        # particles: total conserved, and my slab's particles match the single-domain run's particles in my region
        for sref, sloc in zip(ref.species_list, sim.species_list):
            tot = torch.tensor([sloc.np], dtype=torch.int64, device=dev)
            dist.all_reduce(tot)
            if int(tot) != sref.np:
                ok = False
                print(f"[{rank}] step {step}: particle count {int(tot)} != {sref.np}")
            mine_ref = voxel_table(to_local(sref.particles_host(), nyl, rank, nx, ny, nz), gl.nv)
            mine = voxel_table(sloc.particles_host(), gl.nv)
            if len(mine) != len(mine_ref):
                # a particle within rounding of a slab face may sit on either side for one step; tolerate a handful
                worst["count"] = max(worst["count"], abs(len(mine) - len(mine_ref)))
                continue
            same_vox = np.mean(mine["i"] == mine_ref["i"])
            if same_vox < 0.999:
                ok = False
                print(f"[{rank}] step {step} {sloc.name}: only {same_vox:.5f} of particles in the same voxel")
            sel = mine["i"] == mine_ref["i"]
            for k in ("dx", "dy", "dz", "ux", "uy", "uz"):
                d = np.abs(mine[k][sel] - mine_ref[k][sel]).max() if sel.any() else 0.0
                worst["part"] = max(worst["part"], float(d))
        # fields: my slab's interior nodes vs the same nodes of the single-domain run
        fl = sim.field_array.f.cpu().numpy().reshape(nz + 2, nyl + 2, nx + 2, 20)
        fg = ref.field_array.f.cpu().numpy().reshape(nz + 2, ny + 2, nx + 2, 20)
        a = fl[1:nz + 1, 1:nyl + 1, 1:nx + 1, :16]
        b = fg[1:nz + 1, 1 + rank * nyl:1 + (rank + 1) * nyl, 1:nx + 1, :16]
        for lo, hi in ((0, 3), (4, 7), (12, 15)):     # e, cb, jf
            scale = max(np.abs(b[..., lo:hi]).max(), 1e-12)
            worst["field"] = max(worst["field"], float(np.abs(a[..., lo:hi] - b[..., lo:hi]).max() / scale))
        # energies: sum over ranks vs single domain
        en = torch.tensor(sim.energies(), dtype=torch.float64, device=dev)
        dist.all_reduce(en)
        en_ref = np.array(ref.energies())
        rel = np.abs(en.cpu().numpy() - en_ref) / np.maximum(np.abs(en_ref), 1e-300)
        worst["energy"] = max(worst["energy"], float(rel[6:].max()), float(rel[:6][en_ref[:6] > 1e-12].max(initial=0.0)))

    # tolerances: fp32 accumulation order differs between topologies; values grow slowly over the steps
    ok &= worst["count"] <= 4 and worst["part"] < 5e-4 and worst["field"] < 2e-3 and worst["energy"] < 1e-4
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print(f"multi_gpu_check world={world} steps={args.steps}: worst {worst} -> {'PASS' if int(flag) == 0 else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 0 else 1)


if __name__ == "__main__":
    main()

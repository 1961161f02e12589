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


def to_local(parts, n_local, rank, nx, ny, nz, axis=1):
    """Particles of the global box that live in this rank's slab along `axis`, re-indexed to local voxels."""
    i = parts["i"].astype(np.int64)
    c = [i % (nx + 2), (i // (nx + 2)) % (ny + 2), i // ((nx + 2) * (ny + 2))]
    sel = (c[axis] - 1) // n_local == rank
    out = parts[sel].copy()
    c = [v[sel] for v in c]
    c[axis] = c[axis] - rank * n_local
    ln = [nx, ny, nz]
    ln[axis] = n_local
    out["i"] = (c[0] + (ln[0] + 2) * (c[1] + (ln[1] + 2) * c[2])).astype(np.int32)
    return out


def by_tag(p):
    """Order-free view: every particle carries a unique weight (set at load time, never changed by the push), so
    sorting by it pairs each particle of the slab run with the same particle of the single-domain run."""
    return p[np.argsort(p["w"], kind="stable")]


def tag_weights(parts, base):
    """Unique, exactly representable weights: base * (1 + k * 2^-22), k < 2^17."""
    k = np.arange(len(parts), dtype=np.float64)
    assert len(parts) < (1 << 17)
    parts["w"] = (base * (1.0 + k * 2.0 ** -22)).astype(np.float32)
    assert len(np.unique(parts["w"])) == len(parts)
    return parts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--nx", type=int, default=12)
    ap.add_argument("--ny-per-rank", type=int, default=6)
    ap.add_argument("--nz", type=int, default=8)
    ap.add_argument("--ppc", type=int, default=24)
    ap.add_argument("--axis", type=int, default=1, help="slab axis (0 = x as sample/reconnection, 1 = y as sample/harris)")
    ap.add_argument("--clean", action="store_true",
                    help="divergence cleaning and shared-face synchronisation at short intervals (advance.cc:138-176)")
    ap.add_argument("--harris", action="store_true",
                    help="C4-like: conducting (pec) z walls that reflect particles, sheared B field, drifting species")
    args = ap.parse_args()

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    ax = args.axis
    nloc = args.ny_per_rank                                   # cells per rank along the slab axis
    gn = [args.nx, args.nx, args.nz]                          # global cells; the slab axis gets nloc * world
    gn[ax] = nloc * world
    gn[2] = args.nz
    if ax != 1:
        gn[1] = max(4, args.nx // 2)
    nx, ny, nz = gn
    ln = list(gn); ln[ax] = nloc
    topo = [1, 1, 1]; topo[ax] = world
    dt = G.courant_dt(1, 1, 1, nx, ny, nz, frac=0.97)
    rng = np.random.default_rng(99)
    species = [("electron", -1.0, 1.0, 0.35, (0.0, 0.2, 0.0)), ("ion", 1.0, 4.0, 0.12, (0.0, -0.05, 0.0))]
    if not args.harris:
        species = [(n_, q, m, u, (0, 0, 0)) for n_, q, m, u, _ in species]
    npart = nx * ny * nz * args.ppc
    import refvpic as R
    loads = [tag_weights(R.random_particles(rng, npart, nx, ny, nz, uth=uth, w=1.0, drift=dr), 1.0 / args.ppc)
             for _, _, _, uth, dr in species]

    def setup_walls(g):
        if args.harris:
            for f in (2, 5):
                g.set_fbc(f, G.PEC_FIELDS)
                g.set_pbc(f, G.REFLECT_PARTICLES)

    def initial_fields(nx_, ny_, nz_, x_off=0):
        f = np.zeros(((nx_ + 2) * (ny_ + 2) * (nz_ + 2), 20), np.float32)
        if args.harris:
            z = np.arange(nz_ + 2, dtype=np.float64) - 0.5 * (nz + 1)
            bx = 0.3 * np.tanh(z / 2.0)
            f3 = f.reshape(nz_ + 2, ny_ + 2, nx_ + 2, 20)
            f3[..., 4] = bx[:, None, None].astype(np.float32)            # cbx(z), uniform in x and y
        return f

    # --- single-domain run (every rank has its own copy) ---
    gg = G.partition_periodic_box(0, 0, 0, nx, ny, nz, nx, ny, nz, 1, 1, 1, dt=dt)
    setup_walls(gg)
    dgg = E.DeviceGrid(gg, dev)
    ref = S.Simulation(dgg)
    ref.field_array.f.copy_(torch.from_numpy(initial_fields(nx, ny, nz)))
    for (name, q, m, _, _), load in zip(species, loads):
        sp = ref.define_species(name, q, m, npart, npart, sort_interval=5)
        sp.set_particles(load)
    ref.initialize()

    def maintenance(s):
        if args.clean:
            s.clean_div_e_interval, s.clean_div_b_interval, s.sync_shared_interval = 4, 3, 5

    maintenance(ref)

    # --- slab run ---
    gl = G.partition_periodic_box(0, 0, 0, nx, ny, nz, nx, ny, nz, topo[0], topo[1], topo[2], rank=rank, dt=dt)
    setup_walls(gl)
    dgl = E.DeviceGrid(gl, dev)
    sim = S.Simulation(dgl, exchange=parallel.SlabExchange(dgl, axis=ax))
    sim.field_array.f.copy_(torch.from_numpy(initial_fields(ln[0], ln[1], ln[2])))
    for (name, q, m, _, _), load in zip(species, loads):
        mine = to_local(load, nloc, rank, nx, ny, nz, ax)
        sp = sim.define_species(name, q, m, int(npart / world * 1.6) + 64, npart, sort_interval=5)
        sp.set_particles(mine)
    sim.initialize()
    maintenance(sim)

    worst = dict(count=0, part=0.0, field=0.0, energy=0.0)
    ok = True
    # Outlier bookkeeping.  A particle may deviate by more than rounding only through the one documented mechanism: it
    # landed within rounding of a cell face in one of the two runs (the interpolation of the normal E component is
    # discontinuous there, so the two filings are kicked differently from then on).  Every outlier must be IDENTIFIED
    # as such: its tag must have been seen within 4 ulp of a face, in either run, during the last few steps before it
    # first deviated.  An outlier without that history fails the check.
    near_face_age = [dict() for _ in species]                 # per species: tag -> steps since it was last seen on a face
    known_outliers = [set() for _ in species]
    unexplained = 0
    for step in range(args.steps):
        ref.advance()
        sim.advance()
        sim.sync_counts()                      # the fixed-capacity exchange defers sp.np to the next step
        # particles: total conserved, and my slab's particles match the single-domain run's particles in my region
        for sref, sloc in zip(ref.species_list, sim.species_list):
            tot = torch.tensor([sloc.np], dtype=torch.int64, device=dev)
            dist.all_reduce(tot)
            if int(tot) != sref.np:
                ok = False
                print(f"[{rank}] step {step}: particle count {int(tot)} != {sref.np}")
            mine_ref = by_tag(to_local(sref.particles_host(), nloc, rank, nx, ny, nz, ax))
            mine = by_tag(sloc.particles_host())
            if len(mine) != len(mine_ref) or not np.array_equal(mine["w"], mine_ref["w"]):
                # a particle within rounding of a slab face may sit on either side for one step; tolerate a handful
                common = np.intersect1d(mine["w"], mine_ref["w"])
                worst["count"] = max(worst["count"], max(len(mine), len(mine_ref)) - len(common))
                mine = mine[np.isin(mine["w"], common)]
                mine_ref = mine_ref[np.isin(mine_ref["w"], common)]
            same_vox = np.mean(mine["i"] == mine_ref["i"])
            if same_vox < 0.999:
                ok = False
                print(f"[{rank}] step {step} {sloc.name}: only {same_vox:.5f} of particles in the same voxel")
            sel = mine["i"] == mine_ref["i"]
            # Two runs of the same problem are not bit-reproducible (atomic deposit order), and the reference's
            # interpolation is discontinuous across cell faces in the normal E component: a particle that lands within
            # rounding of a face ends up on either side and is kicked differently from then on.  Such a particle shows up
            # as the same two alternative states in slab and single-domain runs alike (they swap between repetitions),
            # so a handful of outliers per species is run-to-run noise, not a decomposition error.
            delta = np.zeros(len(mine))
            for k in ("dx", "dy", "dz", "ux", "uy", "uz"):
                delta = np.maximum(delta, np.abs(mine[k] - mine_ref[k]) * sel)
            # identify every outlier (deviation beyond fp32 reordering noise, or filed under another voxel)
            si = [sref.name for sref in ref.species_list].index(sref.name)
            ages = near_face_age[si]
            for k_ in list(ages):
                ages[k_] += 1
                if ages[k_] > 6:
                    del ages[k_]
            for arr in (mine, mine_ref):
                edge = np.maximum(np.maximum(np.abs(arr["dx"]), np.abs(arr["dy"])), np.abs(arr["dz"])) > 1.0 - 5e-7
                for t_ in arr["w"][edge]:
                    ages[float(t_)] = 0
            out_mask = (delta > 2e-4) | ~sel
            for t_ in mine["w"][out_mask]:
                t_ = float(t_)
                if t_ in known_outliers[si]:
                    continue
                if t_ in ages:
                    known_outliers[si].add(t_)
                else:
                    unexplained += 1
                    if unexplained <= 5:
                        j = int(np.where(mine["w"] == np.float32(t_))[0][0])
                        print(f"[{rank}] step {step} {sloc.name}: UNEXPLAINED outlier: slab {mine[j]} single {mine_ref[j]}")
            if len(delta) > 8:
                srt = np.sort(delta)
                worst["part"] = max(worst["part"], float(srt[-9]))          # all but the 8 largest
                worst["part_outliers"] = max(worst.get("part_outliers", 0), int((delta > 5e-4).sum()))
                if srt[-1] > 2e-5 and os.environ.get("VPB_CHECK_VERBOSE") and not worst.get("reported"):
                    j = int(np.argmax(delta))
                    print(f"[{rank}] step {step} {sloc.name}: largest particle deviation {srt[-1]:.3e}: slab {mine[j]} single {mine_ref[j]}")
                    worst["reported"] = 1
        # fields: my slab's interior nodes vs the same nodes of the single-domain run
        fl = sim.field_array.f.cpu().numpy().reshape(ln[2] + 2, ln[1] + 2, ln[0] + 2, 20)
        fg = ref.field_array.f.cpu().numpy().reshape(nz + 2, ny + 2, nx + 2, 20)
        a = fl[1:ln[2] + 1, 1:ln[1] + 1, 1:ln[0] + 1, :16]
        sl = [slice(1, nx + 1), slice(1, ny + 1), slice(1, nz + 1)]
        sl[ax] = slice(1 + rank * nloc, 1 + (rank + 1) * nloc)
        b = fg[sl[2], sl[1], sl[0], :16]
        for lo, hi in ((0, 3), (4, 7), (12, 15)):     # e, cb, jf
            scale = max(np.abs(b[..., lo:hi]).max(), 1e-12)
            rel = (np.abs(a[..., lo:hi] - b[..., lo:hi]) / scale).ravel()
            # the nodes around an outlier particle (see above) carry its alternative current: ignore the top 2 %
            worst["field"] = max(worst["field"], float(np.sort(rel)[int(0.98 * (len(rel) - 1))]))
            worst["field_max"] = max(worst.get("field_max", 0.0), float(rel.max()))
        # energies: sum over ranks vs single domain
        en = torch.tensor(sim.energies(), dtype=torch.float64, device=dev)
        dist.all_reduce(en)
        en_ref = np.array(ref.energies())
        rel = np.abs(en.cpu().numpy() - en_ref) / np.maximum(np.abs(en_ref), 1e-300)
        worst["energy"] = max(worst["energy"], float(rel[6:].max()), float(rel[:6][en_ref[:6] > 1e-12].max(initial=0.0)))

    if args.clean:
        # the rms divergence errors and the desynchronisation error the reference prints: same events, same values
        a, b = ref.cleaning_log, sim.cleaning_log
        if [(s_, w_) for s_, w_, _ in a] != [(s_, w_) for s_, w_, _ in b] or not a:
            ok = False
            print(f"[{rank}] cleaning events differ: {len(a)} vs {len(b)}")
        else:
            if rank == 0 and os.environ.get("VPB_CHECK_VERBOSE"):
                for (s_, w_, va), (_, _, vb) in zip(a, b):
                    print(f"  step {s_:3d} {w_:18s} single-domain {va:.6e}  slabs {vb:.6e}")
            for (s_, w_, va), (_, _, vb) in zip(a, b):
                if w_ == "desynchronization":
                    continue                      # measures cross-rank drift: zero on one domain by construction
                rel = abs(va - vb) / max(abs(va), 1e-30)
                if w_.startswith("div_b"):
                    # B starts divergence-free, so this is rounding noise (~1e-9): same magnitude is all that is defined
                    worst["clean_b"] = max(worst.get("clean_b", 0.0), rel)
                else:
                    worst["clean"] = max(worst.get("clean", 0.0), rel)
            ok &= worst.get("clean", 0.0) < 1e-4 and worst.get("clean_b", 0.0) < 0.5
    # tolerances: fp32 accumulation order differs between topologies; values grow slowly over the steps
    worst.pop("reported", None)
    ok &= worst["count"] <= 4 and worst["part"] < 5e-4 and worst["field"] < 2e-3 and worst["energy"] < 1e-4
    ok &= worst.get("part_outliers", 0) <= 8 and worst.get("field_max", 0.0) < 2e-2
    worst["identified_face_bifurcations"] = sum(len(k_) for k_ in known_outliers)
    worst["unexplained_outliers"] = unexplained
    ok &= unexplained == 0
    if not ok:
        print(f"[{rank}] FAILED with worst {worst}")
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print(f"multi_gpu_check world={world} axis={ax} harris={args.harris} steps={args.steps}: worst {worst} -> "
              f"{'PASS' if int(flag) == 0 else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 0 else 1)


if __name__ == "__main__":
    main()

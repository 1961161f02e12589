#!/usr/bin/env python
"""bench.py — particle pushes/sec of the hot path (advance_p + deposit and its glue) on B200.

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU path on the host cores

Workload (BASELINE.json configs[1]): 3-D uniform thermal electron-ion plasma, 128^3 cells per GPU, 64 ppc/species,
periodic, sort_p every 20 steps; synthetic particles, fields start at zero.  A "step" is one pass of the hot-path
block of vpic_simulation::advance (vpic_b200/simulation.py) over every particle.  N > 1: weak scaling, one slab of
128^3 cells per GPU (1 x N x 1 decomposition) with particle migration and halo exchange over NCCL.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
def build_sim(args, rank, world, device):
    import torch
    from vpic_b200 import engine as E, grid as G, simulation as S
    n = args.grid
    nx, ny, nz = n, n, n
    if args.scaling == "strong":                       # fixed global box, y split over the ranks
        if n % world:
            raise SystemExit("--scaling strong needs grid divisible by the number of GPUs")
        ny = n // world
    dt = G.courant_dt(1.0, 1.0, 1.0, nx, ny * world, nz, frac=0.99)
    g = G.partition_periodic_box(0, 0, 0, nx, ny * world, nz, nx, ny * world, nz, 1, world, 1, rank=rank, dt=dt)
    harris = args.workload == "harris"
    if harris:                                         # sample/harris: conducting z walls that reflect particles
        for f in (2, 5):
            g.set_fbc(f, G.PEC_FIELDS)
            g.set_pbc(f, G.REFLECT_PARTICLES)
    dg = E.DeviceGrid(g, device)
    exchange = None
    if world > 1:
        from vpic_b200 import parallel
        exchange = parallel.SlabExchange(dg, axis=1)
    sim = S.Simulation(dg, exchange=exchange)
    sim.deposit_variant = args.variant
    npart = nx * ny * nz * args.ppc
    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + rank)
    # Harris sheet (sample/harris:117-131,216-250 scaled to this box): B_x = b0 tanh((z-zc)/L), two thirds of every
    # species in the sech^2 sheet drifting along +-y, one third as uniform background; same particle count as the
    # uniform load, so the sheet cells hold several times the mean ppc and the lobes a third of it.
    L_sheet, b0, drift = nz / 16.0, 0.15, 0.05
    if harris:
        z = torch.arange(nz + 2, device=device, dtype=torch.float64) - 0.5 - 0.5 * nz     # cell-centred distance from zc
        f3 = sim.field_array.f.view(nz + 2, ny + 2, nx + 2, -1)
        f3[..., 4] = (b0 * torch.tanh(z / L_sheet)).to(torch.float32)[:, None, None]       # cbx(z)
    for k, (name, q, m, uth) in enumerate((("electron", -1.0, 1.0, args.uth), ("ion", 1.0, 25.0, args.uth / 5.0))):
        max_np = int(npart * (1.25 if world > 1 else 1.0)) + 1024
        sp = sim.define_species(name, q, m, max_np, max(int(npart * 0.05), 1 << 16), sort_intervals(args)[k])
        # synthetic load, in random order like the reference's inject_particle loop; sort_p at step 0 orders it
        p = sp.p[:npart]
        p[:, 0:3] = torch.rand((npart, 3), generator=gen, device=device) * 2 - 1
        ix = torch.randint(1, nx + 1, (npart,), generator=gen, device=device, dtype=torch.int32)
        iy = torch.randint(1, ny + 1, (npart,), generator=gen, device=device, dtype=torch.int32)
        iz = torch.randint(1, nz + 1, (npart,), generator=gen, device=device, dtype=torch.int32)
        p[:, 4:7] = torch.randn((npart, 3), generator=gen, device=device) * uth
        if harris:
            n_sheet = (2 * npart) // 3
            u01 = torch.rand((n_sheet,), generator=gen, device=device, dtype=torch.float64).clamp_(1e-9, 1 - 1e-9)
            zs = (0.5 * nz + L_sheet * torch.atanh(2 * u01 - 1)).clamp_(1e-3, nz - 1e-3)   # sech^2 profile, inside the walls
            cell = torch.floor(zs)
            iz[:n_sheet] = cell.to(torch.int32) + 1
            p[:n_sheet, 2] = (2 * (zs - cell) - 1).to(torch.float32)
            p[:n_sheet, 5] += drift if q > 0 else -drift
            del u01, zs, cell
        sp.p.view(torch.int32)[:npart, 3] = ix + (nx + 2) * (iy + (ny + 2) * iz)
        p[:, 7] = 1.0 / args.ppc
        sp.np = npart
        del ix, iy, iz
    sim.initialize()
    return sim


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vpic_b200 import lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    L = lib.load()
    sim = build_sim(args, rank, world, device)
    np_total_local = sum(sp.np for sp in sim.species_list)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        sim.advance()
    barrier()
    launches0 = L.vpb_launch_count()
    sim.push_events = []
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    pushes = 0
    for _ in range(args.steps):
        sim.sync_counts()                      # multi-GPU: particles the last migration appended (deferred read)
        pushes += sum(sp.np for sp in sim.species_list)
        sim.advance()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = L.vpb_launch_count() - launches0
    push_ms = [a.elapsed_time(b) for a, b, _ in sim.push_events]
    push_np = [n for _, _, n in sim.push_events]
    sim.push_events = None
    if args.verbose and rank == 0:
        print("advance_p ms per launch:", " ".join(f"{x:.2f}" for x in push_ms), file=sys.stderr)
    t = torch.tensor([ms, float(pushes)], dtype=torch.float64, device=device)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, pushes = float(tmax[0]), float(tsum[1])
    value = pushes / (ms * 1e-3)

    # roofline of the dominant kernel (advance_p): algorithmic bytes per push = 64 + 176/ppc
    peak, peak_src = peaks()
    bytes_per_push = 64.0 + 176.0 / args.ppc
    avg_push_ms = float(np.mean(push_ms))
    achieved = bytes_per_push * float(np.mean(push_np)) / (avg_push_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "advance_p_kernel", "achieved": round(achieved, 1), "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": load_traffic(), "algorithmic_bytes_per_push": bytes_per_push,
                "avg_launch_ms": round(avg_push_ms, 4), "share_of_step": round(sum(push_ms) / ms, 4)}

    out = None
    if rank == 0:
        out = {"metric": "particle pushes/sec (advance_p+deposit)", "value": value, "unit": "pushes/s",
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": workload_config(args, world, np_total_local), "deposit_variant": args.variant,
               "roofline": roofline, "gpu_launches": int(launches), "clocks": clocks}
    if args.e2e and world == 1:
        e2e = run_e2e(args, device)
        if out is not None:
            out["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, budget_s=args.cpu_seconds)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def load_traffic():
    """dram bytes per advance_p launch from the committed ncu capture (profiles/), if any."""
    p = os.path.join(ROOT, "profiles", "advance_p_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))["dram_bytes_per_launch"]
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------------------------
def run_e2e(args, device):
    """Same step through the reference-facing drop-in symbols (advance_p(species_t*, ...), sort_p, ...) on HOST
    arrays, the way an unmodified reference host program drives them.  Three legs over the same host memory:

      auto      VPB_MODE_AUTO (the library's default): the host may touch any array at any time; arrays the host
                does not touch stay in HBM (page-protection tracking, csrc/lazy_pages.h).  Timed: `--e2e-steps` whole
                steps, each followed by the host reading the kinetic energy of every species and the six field
                energies (dump_energies), then one full host read of both particle arrays (a particle dump).
      coherent  VPB_MODE_COHERENT: every call copies its inputs in and its outputs back (2 timed steps).
      resident  VPB_MODE_RESIDENT: explicit syncs only.
    The headline value is the auto leg; bytes per step are what actually crossed PCIe inside the timed region."""
    import torch
    from vpic_b200 import abi, grid as G, lib
    L = lib.load()
    n = args.grid
    nx = ny = nz = n
    dt = G.courant_dt(1.0, 1.0, 1.0, nx, ny, nz, frac=0.99)
    g = G.partition_periodic_box(0, 0, 0, nx, ny, nz, nx, ny, nz, 1, 1, 1, dt=dt)
    H = HostWorld(L, g, pinned="register")
    npart = nx * ny * nz * args.ppc
    rng = np.random.default_rng(7)
    species = []
    for k, (name, q, m, uth) in enumerate((("electron", -1.0, 1.0, args.uth), ("ion", 1.0, 25.0, args.uth / 5.0))):
        sp = H.new_species(name, q, m, npart, max(int(npart * 0.05), 1 << 16), sort_intervals(args)[k])
        H.fill_uniform(sp, npart, rng, uth, 1.0 / args.ppc)
        species.append(sp)
    en_f = (C.c_double * 6)()
    L.vpic_b200_energy_f.argtypes = [C.c_void_p, C.c_void_p]; L.vpic_b200_energy_f.restype = None

    def timed(mode, steps, diagnostics):
        L.vpic_b200_set_mode(mode)
        t0 = time.perf_counter()
        H.load_interpolator()
        H.advance(species)                                    # first step in a mode: mirrors are (re)filled from the host
        torch.cuda.synchronize()
        first = time.perf_counter() - t0
        tb0 = H.transfer_bytes()
        t0 = time.perf_counter()
        pushes = 0
        energies = None
        movers = 0
        for k in range(steps):
            pushes += sum(sp.c.np for sp in species)
            H.advance(species)
            movers += sum(sp.c.nm for sp in species)          # the per-step result the host reads back: sp->nm
            if diagnostics and (k + 1) % args.e2e_energy_interval == 0:
                # dump_energies at the deck's energies_interval: six field energies and one kinetic energy per species
                L.vpic_b200_energy_f(en_f, C.byref(H.fa))
                energies = [float(x) for x in en_f] + [L.energy_p(C.byref(sp.c), C.byref(H.ia)) for sp in species]
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        tb1 = H.transfer_bytes()
        return {"value": pushes / dt_s, "unit": "pushes/s", "steps": steps, "ms_per_step": 1e3 * dt_s / steps,
                "h2d_bytes_per_step": int((tb1[0] - tb0[0]) / steps), "d2h_bytes_per_step": int((tb1[1] - tb0[1]) / steps),
                "first_step_s": first}, energies

    steps = max(args.e2e_steps, 1)
    auto, energies = timed(2, steps, True)
    # a particle dump in auto mode: the host walks both particle arrays; every device-owned chunk faults back once
    st0 = (C.c_uint64 * 4)(); L.vpic_b200_lazy_stats(st0)
    t0 = time.perf_counter()
    checksum = 0.0
    for sp in species:
        checksum += float(sp.p[:sp.c.np, 7].sum(dtype=np.float64))
    t_dump = time.perf_counter() - t0
    st1 = (C.c_uint64 * 4)(); L.vpic_b200_lazy_stats(st1)
    expect = sum(sp.c.np for sp in species) / args.ppc
    auto["host_reads_all_particles"] = {"seconds": t_dump, "faults": int(st1[0] - st0[0]), "bytes": int(st1[1] - st0[1]),
                                        "weight_checksum_rel_err": abs(checksum - expect) / expect}
    t0 = time.perf_counter()
    H.advance(species)                                        # the step after the dump uploads what the host took back
    torch.cuda.synchronize()
    auto["step_after_dump_s"] = time.perf_counter() - t0
    auto["energies_last_step"] = energies
    coherent, _ = timed(0, 2, False)
    resident, _ = timed(1, max(steps, 10), True)
    L.vpic_b200_set_mode(0)
    auto.update({
        "coherent_mode": coherent, "resident_mode": resident,
        "api": "drop-in extern C symbols (advance_p(species_t*,accumulator_array_t*,interpolator_array_t*), sort_p, "
               "clear/reduce/unload_accumulator_array, field kernels, load_interpolator_array, energy_p, energy_f) on "
               "page-locked HOST arrays, VPB_MODE_AUTO: host memory stays the program's truth (any host access faults "
               f"the data back); every step the host reads sp->nm of both species, every {args.e2e_energy_interval} steps "
               "(a deck's energies_interval) both kinetic energies and the six field energies"})
    return auto


class HostWorld:
    """Host-side reference structs (vpic_b200/abi.py) over pinned numpy arrays, driven through the drop-in symbols."""

    def __init__(self, L, g, pinned=True):
        import torch
        from vpic_b200 import abi
        self.L, self.g, self.abi, self.torch = L, g, abi, torch
        self.keep = []
        self.pinned = pinned if pinned else True
        G = abi.Grid()
        for k in ("dt", "cvac", "eps0", "x0", "y0", "z0", "x1", "y1", "z1", "nx", "ny", "nz", "dx", "dy", "dz", "dV",
                  "rdx", "rdy", "rdz", "r8V"):
            setattr(G, k, getattr(g, k))
        G.sx, G.sy, G.sz, G.nv = 1, g.nx + 2, (g.nx + 2) * (g.ny + 2), g.nv
        for i, b in enumerate(g.bc):
            G.bc[i] = b
        self.range_arr = np.ascontiguousarray(g.range)
        self.nb = self.host_array((g.nv, 6), np.int64, pinned)
        self.nb[:] = g.neighbor
        G.range = self.range_arr.ctypes.data_as(C.POINTER(C.c_int64))
        G.neighbor = self.nb.ctypes.data_as(C.POINTER(C.c_int64))
        G.rangel, G.rangeh = g.rangel, g.rangeh
        self.G = G
        self.fields = self.host_array((g.nv, 20), np.float32, pinned)
        self.interp = self.host_array((g.nv, 20), np.float32, pinned)
        stride = (g.nv + 1) // 2 * 2
        self.accum = self.host_array((stride, 12), np.float32, pinned)
        self.mc = abi.MaterialCoefficient(1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1)
        self.prm = abi.SfaParams(C.pointer(self.mc), 1, 0.0)
        self.fa = abi.FieldArray(self.fields.ctypes.data_as(C.POINTER(C.c_float)), C.pointer(G),
                                 C.cast(C.pointer(self.prm), C.c_void_p))
        self.ia = abi.InterpolatorArray(self.interp.ctypes.data_as(C.POINTER(C.c_float)), C.pointer(G))
        self.aa = abi.AccumulatorArray(self.accum.ctypes.data_as(C.POINTER(C.c_float)), 0, stride, C.pointer(G))
        for fn, argt in (("advance_p", [C.c_void_p] * 3), ("sort_p", [C.c_void_p]), ("load_interpolator_array", [C.c_void_p] * 2),
                         ("clear_accumulator_array", [C.c_void_p]), ("reduce_accumulator_array", [C.c_void_p]),
                         ("unload_accumulator_array", [C.c_void_p] * 2), ("vpic_b200_clear_jf", [C.c_void_p]),
                         ("vpic_b200_synchronize_jf", [C.c_void_p]), ("vpic_b200_advance_b", [C.c_void_p, C.c_float]),
                         ("vpic_b200_advance_e", [C.c_void_p, C.c_float]), ("vpic_b200_transfer_bytes", [C.c_void_p]),
                         ("vpic_b200_set_mode", [C.c_int])):
            f = getattr(L, fn); f.argtypes = argt; f.restype = None
        L.energy_p.restype = C.c_double
        L.energy_p.argtypes = [C.c_void_p, C.c_void_p]

    def host_array(self, shape, dtype, pinned):
        if pinned == "register":
            # What a reference host program has: ordinary anonymous memory, 128-byte aligned (MALLOC_ALIGNED,
            # src/util/util_base.h:150-170) rather than page aligned, page-locked after the fact with cudaHostRegister
            # (what VPIC_B200_PIN=1 does inside the library).  Unlike cudaHostAlloc memory it can be mprotect'ed, which
            # VPB_MODE_AUTO relies on.
            import mmap
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            length = (nbytes + 128 + 4095) // 4096 * 4096
            m = mmap.mmap(-1, length)
            raw = np.frombuffer(m, dtype=np.uint8)
            rc = self.torch.cuda.cudart().cudaHostRegister(raw.ctypes.data, length, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed: {rc}")
            self.keep.append((m, raw))
            return raw[128:128 + nbytes].view(dtype).reshape(shape)
        if pinned:
            t = self.torch.empty(shape, dtype={np.float32: self.torch.float32, np.int64: self.torch.int64,
                                               np.int32: self.torch.int32}[dtype]).pin_memory()
            self.keep.append(t)
            a = t.numpy()
            a[...] = 0
            return a
        return np.zeros(shape, dtype)

    def new_species(self, name, q, m, max_np, max_nm, sort_interval):
        abi = self.abi
        p = self.host_array((max_np, 8), np.float32, self.pinned)
        pm = self.host_array((max_nm, 4), np.float32, self.pinned)
        part = self.host_array((self.g.nv + 1,), np.int32, self.pinned)
        sp = abi.Species()
        sp.name = name.encode(); sp.q, sp.m = q, m
        sp.np, sp.max_np, sp.nm, sp.max_nm = 0, max_np, 0, max_nm
        sp.p = p.ctypes.data_as(C.POINTER(abi.Particle)); sp.pm = pm.ctypes.data_as(C.POINTER(abi.ParticleMover))
        sp.last_sorted = -(2 ** 63); sp.sort_interval = sort_interval; sp.sort_out_of_place = 0
        sp.partition = part.ctypes.data_as(C.POINTER(C.c_int32)); sp.g = C.pointer(self.G)

        class W:
            pass
        w = W(); w.c = sp; w.p = p; w.pm = pm; w.partition = part
        return w

    def fill_uniform(self, sp, n, rng, uth, w):
        g = self.g
        t = self.torch
        gen = t.Generator().manual_seed(int(rng.integers(1 << 31)))
        p = t.from_numpy(sp.p)[:n]
        p[:, 0:3] = t.rand((n, 3), generator=gen) * 2 - 1
        ix = t.randint(1, g.nx + 1, (n,), generator=gen, dtype=t.int32)
        iy = t.randint(1, g.ny + 1, (n,), generator=gen, dtype=t.int32)
        iz = t.randint(1, g.nz + 1, (n,), generator=gen, dtype=t.int32)
        t.from_numpy(sp.p).view(t.int32)[:n, 3] = ix + (g.nx + 2) * (iy + (g.ny + 2) * iz)
        p[:, 4:7] = t.randn((n, 3), generator=gen) * uth
        p[:, 7] = w
        sp.c.np = n

    def load_interpolator(self):
        self.L.load_interpolator_array(C.byref(self.ia), C.byref(self.fa))

    def transfer_bytes(self):
        out = (C.c_uint64 * 2)()
        self.L.vpic_b200_transfer_bytes(out)
        return int(out[0]), int(out[1])

    def advance(self, species):
        L, fa, ia, aa = self.L, C.byref(self.fa), C.byref(self.ia), C.byref(self.aa)
        step = self.G.step
        for sp in species:
            if step % sp.c.sort_interval == 0:
                L.sort_p(C.byref(sp.c))
        L.clear_accumulator_array(aa)
        for sp in species:
            L.advance_p(C.byref(sp.c), aa, ia)
        L.reduce_accumulator_array(aa)
        L.vpic_b200_clear_jf(fa)
        L.unload_accumulator_array(fa, aa)
        L.vpic_b200_synchronize_jf(fa)
        L.vpic_b200_advance_b(fa, 0.5)
        L.vpic_b200_advance_e(fa, 1.0)
        L.vpic_b200_advance_b(fa, 0.5)
        L.load_interpolator_array(ia, fa)
        self.G.step += 1


# ------------------------------------------------------------------------------------------------------------------
def ref_variants():
    """Reference builds this host CPU can run, strongest instruction set first: V16 AVX-512 (CMakeLists.txt:79), V8
    AVX2+FMA, V4 SSE, scalar.  Which one is FASTEST on this host is measured, not assumed (pick_ref_variant)."""
    flags = ""
    try:
        flags = open("/proc/cpuinfo").read()
    except Exception:
        pass
    order = []
    if all(f" {x}" in flags for x in ("avx512f", "avx512dq", "avx512bw", "avx512vl", "avx2", "fma")):
        order.append("v16")
    if " avx2" in flags and " fma" in flags:
        order.append("v8")
    order += ["v4", "scalar"]
    return [v for v in order if os.path.exists(os.path.join(ROOT, "oracle", "_ref", f"libvpic_ref_{v}.so"))]


_picked = {}


def pick_ref_variant(args=None):
    """The fastest runnable reference build: every SIMD candidate is timed on a 32^3-cell probe of the workload in its
    own subprocess (one reference build can be booted per process) and the best one wins."""
    if "v" in _picked:
        return _picked["v"]
    cands = ref_variants()
    best, rates = (cands[0] if cands else None), {}
    simd = [v for v in cands if v in ("v16", "v8")]
    if len(simd) > 1 and args is not None:
        for v in simd:
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--probe-variant", v,
                                      "--ppc", str(args.ppc), "--uth", str(args.uth), "--sort-interval", str(args.sort_interval)],
                                     capture_output=True, text=True, timeout=300)
                rates[v] = float(json.loads(out.stdout.strip().splitlines()[-1])["value"])
            except Exception:
                rates[v] = 0.0
        best = max(simd, key=lambda v: rates[v])
    _picked["v"], _picked["rates"] = best, rates
    return best


def fill_reference_species(sp, npart, nx, ny, nz, uth, w, rng, chunk=1 << 24):
    """The synthetic load of build_sim written straight into the reference's particle array, a chunk at a time (the
    full workload holds 8.6 GB of particles; no second copy is made)."""
    p = sp.p
    for a in range(0, npart, chunk):
        b = min(npart, a + chunk)
        n = b - a
        v = p[a:b]
        for k in ("dx", "dy", "dz"):
            v[k] = rng.random(n, dtype=np.float32) * np.float32(2) - np.float32(1)
        ix = rng.integers(1, nx + 1, n, dtype=np.int32)
        iy = rng.integers(1, ny + 1, n, dtype=np.int32)
        iz = rng.integers(1, nz + 1, n, dtype=np.int32)
        v["i"] = ix + (nx + 2) * (iy + (ny + 2) * iz)
        for k in ("ux", "uy", "uz"):
            v[k] = rng.standard_normal(n, dtype=np.float32) * np.float32(uth)
        v["w"] = np.float32(w)
    sp.c.np = npart
    sp.c.nm = 0


def reference_sample(args, n_steps, grid_n, warm=1, variant=None):
    """Time the reference's own CPU hot path (unmodified sources, oracle/_ref): the same plasma on a grid_n^3 box."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refvpic as R
    variant = variant or pick_ref_variant(args)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if variant is None:
        return port_sample(args, grid_n)
    lib = R.load_ref(variant, tpp=cores)
    nx = ny = nz = grid_n
    W = R.RefWorld(lib, nx, ny, nz, dt=None)
    W.g.contents.dt = float(np.float32(0.99 / np.sqrt(3.0)))
    npart = nx * ny * nz * args.ppc
    rng = np.random.default_rng(5)
    species = []
    for k, (name, q, m, uth) in enumerate((("electron", -1.0, 1.0, args.uth), ("ion", 1.0, 25.0, args.uth / 5.0))):
        sp = W.new_species(name, q, m, npart, max(int(npart * 0.05), 1 << 16), sort_intervals(args)[k])
        fill_reference_species(sp, npart, nx, ny, nz, uth, 1.0 / args.ppc, rng)
        species.append(sp)
    lib.load_interpolator_array(W.ia, W.fa)

    def step(k):
        for sp, every in zip(species, sort_intervals(args)):
            if k % every == 0:
                lib.sort_p(sp.sp)
        lib.clear_accumulator_array(W.aa)
        for sp in species:
            lib.advance_p(sp.sp, W.aa, W.ia)
        lib.reduce_accumulator_array(W.aa)
        W.clear_jf()
        lib.unload_accumulator_array(W.fa, W.aa)
        W.synchronize_jf()
        W.advance_b(0.5); W.advance_e(1.0); W.advance_b(0.5)
        lib.load_interpolator_array(W.ia, W.fa)

    k = 0
    for _ in range(warm):
        step(k); k += 1
    times = []
    for _ in range(n_steps):
        t0 = time.perf_counter()
        step(k); k += 1
        times.append(time.perf_counter() - t0)
    total = sum(times)
    pushes = 2 * npart * n_steps
    probe = _picked.get("rates") or {}
    return {"value": pushes / total, "unit": "pushes/s", "cores": cores, "kind": "reference",
            "sample": f"unmodified reference ({variant} build, pthreads --tpp {cores}), same plasma on a {grid_n}^3-cell box, "
                      f"{args.ppc} ppc/species, {n_steps} timed steps after {warm} warm-up (first step sorts)"
                      + (f"; fastest of the SIMD builds this CPU runs, probed on 32^3 cells: "
                         + ", ".join(f"{k} {v / 1e6:.0f} M pushes/s" for k, v in probe.items()) if probe else ""),
            "variant": variant, "grid": grid_n, "ms_per_step": 1e3 * total / n_steps, "steps": n_steps}


def port_sample(args, grid_n):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refvpic as R
    from vpic_b200 import grid as G, abi
    orc = R.load_oracle()
    nx = ny = nz = min(grid_n, 32)
    g = G.partition_periodic_box(0, 0, 0, nx, ny, nz, nx, ny, nz, 1, 1, 1, dt=G.courant_dt(1, 1, 1, nx, ny, nz, frac=0.99))
    n = nx * ny * nz * args.ppc
    rng = np.random.default_rng(5)
    parts = R.random_particles(rng, n, nx, ny, nz, uth=args.uth, w=1.0 / args.ppc)
    interp = np.zeros((g.nv, 20), np.float32)
    acc = np.zeros(((g.nv + 1) // 2 * 2, 12), np.float32)
    pm = np.zeros(1024, dtype=abi.mover_dtype)
    f32 = np.float32
    a = R.OraclePushArgs(parts.ctypes.data, n, pm.ctypes.data, 1024, interp.ctypes.data, 20, acc.ctypes.data, 12,
                         g.neighbor.ctypes.data, g.rangel, g.rangeh, f32(-0.5 * g.dt), f32(g.dt), f32(g.dt), f32(g.dt), f32(-1))
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        orc.vpo_advance_p(C.byref(a), None)
    dt_s = time.perf_counter() - t0
    return {"value": n * reps / dt_s, "unit": "pushes/s", "cores": 1, "kind": "port",
            "sample": f"oracle C port, advance_p only, {nx}^3 cells x {args.ppc} ppc, {reps} passes"}


def cpu_baseline(args, budget_s=15.0):
    # one probing step on a small box, then size the sample to ~budget_s of CPU work
    try:
        if pick_ref_variant(args) is None:
            return port_sample(args, 32)
        rates = _picked.get("rates") or {}
        rate = max(rates.values()) if rates and max(rates.values()) > 0 else 5e8
        grid_n = 64
        per_step = 2 * grid_n ** 3 * args.ppc / rate
        n_steps = int(max(2, min(20, budget_s / max(per_step, 1e-3))))
        return reference_sample(args, n_steps, grid_n, warm=1)
    except Exception as e:   # report, never fake
        return {"value": None, "unit": "pushes/s", "cores": None, "kind": "reference", "sample": f"failed: {e!r}"}


def sort_intervals(args):
    """--sort-interval N or E,I: species_t.sort_interval of the electrons and of the ions (a deck sets it per species,
    sample/harris:181-182 uses 40 for the ions and 20 for the electrons)."""
    v = [int(x) for x in str(args.sort_interval).split(",")]
    return (v[0], v[0]) if len(v) == 1 else (v[0], v[1])


def workload_config(args, world, np_total_local):
    """`config` of the JSON line; identical on both arms (the reference arm runs the same workload on the host)."""
    return {"workload": (f"uniform thermal e-/ion plasma, {args.grid}^3 cells " if args.workload == "uniform" else
                         f"Harris current sheet (sech^2 sheet + background, B_x = b0 tanh(z/L), drifting e-/ion, "
                         f"conducting reflecting z walls), {args.grid}^3 cells ")
                        + ("per GPU" if args.scaling == "weak" else "in total")
                        + f", {args.ppc} ppc/species" + (" on average" if args.workload == "harris" else "")
                        + f", periodic{' in x and y' if args.workload == 'harris' else ''}, "
                        f"sort_p every {sort_intervals(args)[0]} (electrons) / {sort_intervals(args)[1]} (ions) steps"
                        + (" (BASELINE.json configs[1])" if (args.grid, args.ppc, args.workload) == (128, 64, "uniform") else ""),
            "particles_per_gpu": np_total_local, "decomposition": f"1x{world}x1 slabs",
            "l2": f"particle arrays ({np_total_local * 32 / 1e9:.1f} GB per GPU) exceed the 126 MB L2; no flush needed"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.probe_variant:                                  # child of pick_ref_variant: one 32^3 probe of one build
        res = reference_sample(args, 2, 32, warm=1, variant=args.probe_variant)
        print(json.dumps({"value": res["value"]}), flush=True)
        return
    # The reference arm runs the SAME configuration as our arm at N=1 (the whole 128^3-cell, 64 ppc, two-species box),
    # every step a whole step of it, on all host threads; --ref-grid shrinks it for a quick look.
    grid_n = args.ref_grid or args.grid
    res = reference_sample(args, args.steps, grid_n, warm=max(1, args.warmup))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cfg = workload_config(args, 1, 2 * grid_n ** 3 * args.ppc)
    if grid_n != args.grid:
        cfg["workload"] += f"; each step a bounded sample of it: {grid_n}^3 cells on the host cores"
    out = {"impl": "reference", "metric": "particle pushes/sec (advance_p+deposit)", "value": res["value"],
           "unit": "pushes/s", "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup),
           "ms_per_step": res.get("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": cfg,
           "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "variant")},
           "e2e": {"value": res["value"], "unit": "pushes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--ppc", type=int, default=64)
    ap.add_argument("--uth", type=float, default=0.18)
    ap.add_argument("--sort-interval", default="6,12",
                    help="steps between sort_p calls (species_t::sort_interval): N, or E,I for electrons and ions.  The "
                         "default is the measured optimum of this build at C2 (profiles/r02_sort_interval_sweep.txt); the "
                         "reference arm runs the same intervals.  C5 names 20.")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--e2e", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=60)
    ap.add_argument("--e2e-energy-interval", type=int, default=10,
                    help="e2e leg: steps between dump_energies-style diagnostics (energy_f + energy_p per species)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-grid", type=int, default=0, help="reference arm: cells per side (default: --grid, the full workload)")
    ap.add_argument("--probe-variant", default="", help=argparse.SUPPRESS)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--workload", default="uniform", choices=["uniform", "harris"],
                    help="uniform thermal plasma (BASELINE.json configs[1], the default) or a Harris current sheet")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): one grid^3 slab per GPU; strong: one grid^3 box split over the GPUs")
    args = ap.parse_args()
    if args.workload == "harris":        # the end-to-end and CPU legs are defined on the uniform workload only
        args.e2e, args.no_cpu_baseline = 0, True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

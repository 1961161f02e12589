"""Host-side logic and the C-ABI surface, without a GPU."""
import ctypes as C
import os
import re
import numpy as np
import pytest

import refvpic as R
from vpic_b200 import grid as G, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    L = lib.load()
    assert L.vpb_version() == 100
    hdr = open(os.path.join(ROOT, "include", "vpic_b200.h")).read()
    declared = set(re.findall(r"\b(vpb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "header parse failed"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/vpic_b200.h but not exported"
    dh = os.path.join(ROOT, "include", "vpic_b200_dropin.h")
    if os.path.exists(dh):
        for name in re.findall(r"^\s*(?:void|int|double)\s+([a-z_0-9]+)\s*\(", open(dh).read(), re.M):
            assert hasattr(L, name), f"{name} declared in include/vpic_b200_dropin.h but not exported"


def test_no_cpu_fallback_without_device():
    """Without a CUDA device a compute call must fail loudly, not silently run elsewhere."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = lib.load()
    n = C.c_int(0)
    rc = L.vpb_device_count(C.byref(n))
    assert rc != 0 or n.value == 0
    p = C.c_void_p()
    rc = L.vpb_malloc(C.byref(p), 1024)
    assert rc != 0 and b"CUDA" in L.vpb_last_error()
    with pytest.raises(lib.VpbError):
        lib.check(rc, "vpb_malloc")


def test_bad_args_are_rejected():
    L = lib.load()
    assert L.vpb_advance_p(None, None) != 0
    assert b"Bad args" in L.vpb_last_error()
    assert L.vpb_sort_p(None, 5, None, None, 4, 4, 4, None, 0, None) != 0
    assert L.vpb_load_interpolator(None, 20, None, 4, 4, 4, None) != 0
    a = lib.FieldArgs()
    assert L.vpb_vacuum_advance_e(C.byref(a), 1.0, None) != 0


def test_index_sort_entry_points_reject_bad_args():
    """The deferred-sort entry points (vpb_sort_p_index, vpb_permute_p, vpb_unpermute_p, vpb_extract_keys, and
    vpb_advance_p with perm) validate their arguments before touching the device."""
    L = lib.load()
    assert L.vpb_sort_p_index(None, None, 5, None, None, 4, 4, 4, None, 0, None, 0, None, None) != 0
    assert b"Bad args" in L.vpb_last_error()
    assert L.vpb_permute_p(None, 5, None, None, None) != 0
    assert L.vpb_unpermute_p(None, 5, None, None, None) != 0
    assert L.vpb_extract_keys(None, 5, None, None) != 0
    assert L.vpb_sort_index_work_bytes(1000) >= 1000 * 20 and L.vpb_sort_index_scratch_bytes(1000, 216) > 0
    a = lib.PushArgs()
    buf = (C.c_char * 4096)()
    base = (C.addressof(buf) + 127) // 128 * 128
    a.p = a.interp = a.accum = a.neighbor = a.counters = a.pm = base
    a.np, a.max_nm, a.interp_stride, a.accum_stride, a.nx, a.ny, a.nz = 4, 4, 20, 12, 2, 2, 2
    a.perm = base                                  # an order without somewhere else to put the particles
    assert L.vpb_advance_p(C.byref(a), None) != 0 and b"p_out" in L.vpb_last_error()
    a.p_out = base                                 # ... or onto themselves
    assert L.vpb_advance_p(C.byref(a), None) != 0 and b"p_out" in L.vpb_last_error()


def test_bench_sort_intervals():
    import types, bench
    assert bench.sort_intervals(types.SimpleNamespace(sort_interval="20")) == (20, 20)
    assert bench.sort_intervals(types.SimpleNamespace(sort_interval="6,12")) == (6, 12)
    assert bench.sort_intervals(types.SimpleNamespace(sort_interval=25)) == (25, 25)


@pytest.mark.parametrize("dims", [(6, 5, 4), (8, 1, 8), (64, 64, 1)])
def test_grid_matches_reference(ref_scalar, dims):
    nx, ny, nz = dims
    W = R.RefWorld(ref_scalar, nx, ny, nz, lx=2.0 * nx, ly=0.5 * ny, lz=1.25 * nz,
                   pbc={0: -1, 5: -2}, fbc={0: -1})
    gc = W.g.contents
    g = G.partition_periodic_box(0, 0, 0, 2.0 * nx, 0.5 * ny, 1.25 * nz, nx, ny, nz, 1, 1, 1, dt=gc.dt)
    g.set_pbc(0, -1); g.set_pbc(5, -2); g.set_fbc(0, -1)
    assert np.array_equal(g.neighbor, W.neighbor)
    for k in ("dx", "dy", "dz", "dV", "rdx", "rdy", "rdz", "r8V", "x0", "x1", "y1", "z1", "dt"):
        assert np.float32(getattr(g, k)) == np.float32(getattr(gc, k)), k
    assert g.rangel == gc.rangel and g.rangeh == gc.rangeh and g.nv == gc.nv
    assert list(g.bc) == list(gc.bc)


def test_slab_grid_neighbours():
    """1 x N x 1 slab decomposition: remote faces carry the neighbour rank's global voxel ids."""
    N, nx, ny, nz = 4, 6, 8, 5
    grids = [G.partition_periodic_box(0, 0, 0, nx, ny, nz, nx, ny, nz, 1, N, 1, rank=r, dt=0.1) for r in range(N)]
    for r, g in enumerate(grids):
        assert (g.nx, g.ny, g.nz) == (nx, ny // N, nz)
        up, dn = grids[(r + 1) % N], grids[(r - 1) % N]
        v = G.voxel(3, g.ny, 2, g.nx, g.ny, g.nz)
        assert g.neighbor[v, 4] == up.rangel + G.voxel(3, 1, 2, g.nx, g.ny, g.nz)
        v = G.voxel(3, 1, 2, g.nx, g.ny, g.nz)
        assert g.neighbor[v, 1] == dn.rangel + G.voxel(3, dn.ny, 2, g.nx, g.ny, g.nz)
        assert g.face_codes() == [0, 1, 0, 0, 1, 0]


def test_lazy_pages_state_machine():
    """VPB_MODE_AUTO's page-protection tracker (vpic_b200/csrc/lazy_pages.cpp) against a fake device: faults bring
    chunks back, sparse host writes survive, edges are copied eagerly, syscalls are served by host_access, concurrent
    faulting threads never see a half-filled chunk, remapped arrays are noticed — and a wild access still crashes."""
    import signal
    import subprocess
    csrc = os.path.join(ROOT, "vpic_b200", "csrc")
    subprocess.check_call(["make", "-s", "-C", csrc, "lazy_test"])
    exe = os.path.join(csrc, "build", "lazy_pages_test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "lazy_pages_test: ok" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([exe, "--crash"], capture_output=True, text=True, timeout=120)
    assert r.returncode == -signal.SIGSEGV, (r.returncode, r.stdout, r.stderr)
    # randomised model check: entry points with random extents, host reads / writes, syscalls, concurrent readers
    for seed in (11, 12, 13):
        r = subprocess.run([exe, "--stress", "4000", str(seed)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "stress ok" in r.stdout, r.stdout + r.stderr
    # the same with the array ends copied back asynchronously (one wait per entry point, then after_sync)
    for seed in (21, 22, 23):
        r = subprocess.run([exe, "--stress-async", "4000", str(seed)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "stress ok" in r.stdout, r.stdout + r.stderr


def test_simulation_step_follows_advance_cc_order(monkeypatch):
    """vpic_b200/simulation.py must call the operators in the order of vpic_simulation::advance
    (src/vpic/advance.cc:25-185), including the divergence-cleaning block at its intervals.  The operators are replaced
    by recorders, so this runs without a GPU."""
    from types import SimpleNamespace as NS
    import torch
    from vpic_b200 import engine as E, simulation as S
    log = []
    g = G.partition_periodic_box(0, 0, 0, 4, 4, 4, 4, 4, 4, 1, 1, 1, dt=0.1)
    dg = E.DeviceGrid(g, device="cpu")

    class FakeFields:
        def __init__(self, *a, **k):
            self._en = torch.zeros(6, dtype=torch.float64)
        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            def call(*a, **k):
                log.append(name if not a else f"{name}({a[0]})")
                return [1.0, 1.0] if name == "rms_terms" else (0.5 if name == "rms_finish" else None)
            return call

    monkeypatch.setattr(E, "FieldArray", FakeFields)
    kw = []                                              # (operator, species, keyword arguments) of the particle operators
    def recorder(name):
        def call(*a, **k):
            log.append(name)
            if name in ("sort_p", "advance_p"):
                kw.append((name, a[0].name, {x: k[x] for x in ("defer", "emit_keys") if x in k}))
        return call
    for fn in ("sort_p", "clear_accumulator_array", "advance_p", "reduce_accumulator_array", "unload_accumulator_array",
               "load_interpolator_array", "accumulate_rho_p", "finish_advance_p_all"):
        monkeypatch.setattr(E, fn, recorder(fn))
    sim = S.Simulation(dg)
    assert sim.defer_sort
    for name in ("e", "i"):
        sim.species_list.append(NS(name=name, sort_interval=2, nm=0, np=10))
    sim.clean_div_e_interval, sim.clean_div_b_interval, sim.sync_shared_interval = 2, 3, 4
    sim.num_div_e_round = sim.num_div_b_round = 2

    def step():
        log.clear()
        sim.advance()
        return list(log)

    push = ["clear_accumulator_array", "advance_p", "advance_p", "finish_advance_p_all", "reduce_accumulator_array"]
    fields = ["clear_jf", "unload_accumulator_array", "synchronize_jf", "advance_b(0.5)", "advance_e(1.0)", "advance_b(0.5)"]
    div_e = ["clear_rhof", "accumulate_rho_p", "accumulate_rho_p", "synchronize_rho",
             "compute_div_e_err", "rms_terms(vpb_compute_rms_div_e_err)", "rms_finish([1.0, 1.0])", "clean_div_e",
             "compute_div_e_err", "rms_terms(vpb_compute_rms_div_e_err)", "rms_finish([1.0, 1.0])", "clean_div_e"]
    div_b = ["compute_div_b_err", "rms_terms(vpb_compute_rms_div_b_err)", "rms_finish([1.0, 1.0])", "clean_div_b",
             "compute_div_b_err", "rms_terms(vpb_compute_rms_div_b_err)", "rms_finish([1.0, 1.0])", "clean_div_b"]
    sync = ["synchronize_tang_e_norm_b"]
    # step 0: both species sort, every maintenance interval divides 0
    assert step() == ["sort_p", "sort_p"] + push + fields + div_e + div_b + sync + ["load_interpolator_array"]
    # step 1: nothing periodic
    assert step() == push + fields + ["load_interpolator_array"]
    # step 2: sort and div-E clean; step 3: div-B clean only
    assert step() == ["sort_p", "sort_p"] + push + fields + div_e + ["load_interpolator_array"]
    assert step() == push + fields + div_b + ["load_interpolator_array"]
    assert sim.step == 4 and [w for _, w, _ in sim.cleaning_log][:3] == ["div_e initial", "div_e cleaned", "div_b initial"]
    # sort_p only computes the order (the push that follows applies it), and the push of the step before a sort is asked
    # to leave the voxel keys behind: sort_interval = 2, so the pushes of steps 1 and 3
    assert all(k == {"defer": True} for n, _, k in kw if n == "sort_p")
    pushes = [k["emit_keys"] for n, _, k in kw if n == "advance_p"]
    assert pushes == [False, False, True, True, False, False, True, True]

"""Pins the C restatement (oracle/liboracle.so) bit-for-bit against the UNMODIFIED reference compiled from
/root/reference (oracle/_ref/libvpic_ref_scalar.so).  CPU only."""
import ctypes as C
import os
import numpy as np
import pytest

import refvpic as R
from vpic_b200 import abi


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def make_world(lib, rng, nx, ny, nz, **kw):
    W = R.RefWorld(lib, nx, ny, nz, **kw)
    W.fields[:] = R.random_fields(rng, W.nv)
    lib.load_interpolator_array(W.ia, W.fa)
    return W


def port_push(orc, W, sp, parts, max_nm):
    p2 = parts.copy()
    pm2 = np.zeros(max_nm, dtype=abi.mover_dtype)
    acc2 = np.zeros((W.aa.contents.stride, W.asf), dtype=np.float32)
    k = sp.push_constants()
    a = R.OraclePushArgs(p2.ctypes.data, len(p2), pm2.ctypes.data, max_nm, W.interp.ctypes.data, W.isf,
                         acc2.ctypes.data, W.asf, W.neighbor.ctypes.data, W.g.contents.rangel, W.g.contents.rangeh,
                         k["qdt_2mc"], k["cdt_dx"], k["cdt_dy"], k["cdt_dz"], k["qsp"])
    ign = C.c_int32(0)
    nm = orc.vpo_advance_p(C.byref(a), C.byref(ign))
    return p2, pm2[:nm], acc2, ign.value


@pytest.mark.parametrize("dims,uth,pbc", [
    ((6, 5, 4), 0.5, None),                       # periodic, many crossings
    ((8, 1, 8), 0.3, None),                       # 2-D deck shape (ny == 1: the +-y neighbour is the voxel itself)
    ((5, 4, 3), 0.6, {0: -1, 3: -1}),             # reflecting x walls
    ((5, 4, 3), 0.6, {2: -2, 5: -2}),             # absorbing z walls -> movers handed to boundary_p
])
def test_advance_p_bit_exact(ref_scalar, oracle, dims, uth, pbc):
    rng = np.random.default_rng(11)
    nx, ny, nz = dims
    W = make_world(ref_scalar, rng, nx, ny, nz, pbc=pbc)
    sp = W.new_species("e%d" % rng.integers(1 << 30), -1.0, 1.0, 4096, 4096)
    n = 3200                                        # multiple of 16: one pipeline block holds every deposit
    parts = R.random_particles(rng, n, nx, ny, nz, uth=uth)
    sp.set_particles(parts)
    ref_scalar.clear_accumulator_array(W.aa)
    ref_scalar.advance_p(sp.sp, W.aa, W.ia)
    ref_scalar.reduce_accumulator_array(W.aa)
    p2, pm2, acc2, ign = port_push(oracle, W, sp, parts, 4096)
    assert sp.c.nm == len(pm2)
    if pbc and -2 in pbc.values():
        assert sp.c.nm > 0
    assert np.array_equal(bits(p2), bits(sp.p[:n]))
    assert np.array_equal(bits(pm2), bits(sp.pm[:sp.c.nm]))
    assert np.array_equal(bits(acc2), bits(W.accum[0]))
    assert (p2["i"] != parts["i"]).mean() > 0.1     # the move_p path really ran


def test_sort_p_bit_exact(ref_scalar, oracle):
    rng = np.random.default_rng(5)
    nx, ny, nz = 7, 6, 5
    W = make_world(ref_scalar, rng, nx, ny, nz)
    sp = W.new_species("s%d" % rng.integers(1 << 30), -1.0, 1.0, 20000, 16)
    parts = R.random_particles(rng, 17777, nx, ny, nz)
    sp.set_particles(parts)
    ref_scalar.sort_p(sp.sp)
    p2, aux = parts.copy(), np.zeros_like(parts)
    part2 = np.full(W.nv + 1, -7, dtype=np.int32)
    oracle.vpo_sort_p(p2.ctypes.data, len(p2), aux.ctypes.data, part2.ctypes.data, nx, ny, nz)
    assert np.array_equal(bits(p2), bits(sp.p[:len(p2)]))
    assert np.array_equal(part2[:W.nv], sp.partition[:W.nv])
    assert np.all(np.diff(p2["i"]) >= 0)


def test_interpolator_unload_energy_center(ref_scalar, oracle):
    rng = np.random.default_rng(9)
    nx, ny, nz = 6, 4, 5
    W = make_world(ref_scalar, rng, nx, ny, nz)
    lib = ref_scalar
    i2 = np.zeros_like(W.interp)
    oracle.vpo_load_interpolator(i2.ctypes.data, W.isf, W.fields.ctypes.data, nx, ny, nz)
    assert np.array_equal(bits(i2), bits(W.interp))

    # unload: random accumulators in the interior, zero ghosts as the reference requires
    acc = W.accum[0]
    acc[:] = 0
    x, y, z = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    v = abi.voxel(x, y, z, nx, ny, nz).ravel()
    acc[v] = rng.normal(0, 1, (len(v), W.asf)).astype(np.float32)
    f2 = W.fields.copy()
    g = W.g.contents
    oracle.vpo_unload_accumulator(f2.ctypes.data, acc.ctypes.data, W.asf, nx, ny, nz, g.rdx, g.rdy, g.rdz, g.dt)
    lib.unload_accumulator_array(W.fa, W.aa)
    assert np.array_equal(bits(f2), bits(W.fields))
    # clear
    a2 = acc.copy()
    oracle.vpo_clear_accumulator(a2.ctypes.data, W.asf, nx, ny, nz)
    lib.clear_accumulator_array(W.aa)
    assert np.array_equal(bits(a2), bits(W.accum[0]))

    sp = W.new_species("c%d" % rng.integers(1 << 30), -1.0, 1.5, 8192, 16)
    parts = R.random_particles(rng, 4800, nx, ny, nz, uth=0.4, w=0.37)
    sp.set_particles(parts)
    e_ref = lib.energy_p(sp.sp, W.ia)
    e_port = oracle.vpo_energy_p(parts.ctypes.data, len(parts), W.interp.ctypes.data, W.isf, -1.0, 1.5, g.dt, g.cvac)
    assert e_ref == e_port
    k = sp.push_constants()
    p2 = parts.copy()
    oracle.vpo_uncenter_p(p2.ctypes.data, len(p2), W.interp.ctypes.data, W.isf, k["qdt_2mc"])
    lib.uncenter_p(sp.sp, W.ia)
    assert np.array_equal(bits(p2), bits(sp.p[:len(p2)]))
    oracle.vpo_center_p(p2.ctypes.data, len(p2), W.interp.ctypes.data, W.isf, k["qdt_2mc"])
    lib.center_p(sp.sp, W.ia)
    assert np.array_equal(bits(p2), bits(sp.p[:len(p2)]))


# one anisotropic, conducting material filling space: the reference still uses its vacuum_* kernels (sfa.cc:202-211)
DIELECTRIC = (1.5, 2.0, 2.5, 1.2, 1.1, 1.3, 0.1, 0.2, 0.05, 0.0, 0.0, 0.0)


@pytest.mark.parametrize("dims,fbc,damp,material", [
    ((6, 5, 4), None, 0.0, None),
    ((6, 5, 4), None, 0.01, None),
    ((8, 8, 1), {0: -1, 3: -1}, 0.0, None),             # harris-like: pec x walls, degenerate z
    ((5, 1, 7), {2: -2, 5: -3}, 0.0, None),             # symmetric / pmc walls, degenerate y
    ((6, 5, 4), {0: -4, 3: -4, 2: -4, 5: -1}, 0.0, None),  # absorbing (Higdon) walls on -x, +x, -z; pec on +z
    ((9, 1, 6), {0: -4, 3: -4}, 0.01, None),            # lpi-like: 2-D, absorbing x walls
    ((6, 5, 4), None, 0.01, DIELECTRIC),
    ((7, 4, 5), {0: -1, 3: -4}, 0.0, DIELECTRIC),
])
def test_field_advance_bit_exact(ref_scalar, oracle, dims, fbc, damp, material):
    rng = np.random.default_rng(3)
    nx, ny, nz = dims
    W = R.RefWorld(ref_scalar, nx, ny, nz, fbc=fbc, damp=damp, material=material)
    if material is not None:
        mc = W.material_coefficients()
        assert mc[0] != 1.0 and mc[1] != 1.0 and mc[6] != 1.0 and mc[10] == 1.5
    f0 = R.random_fields(rng, W.nv)
    f0[:, 8:11] = rng.normal(0, 0.01, (W.nv, 3))       # tca
    f0[:, 12:15] = rng.normal(0, 0.02, (W.nv, 3))      # jf
    W.fields[:] = f0
    f2 = f0.copy()
    a = W.field_args(f2)
    for _ in range(3):
        W.synchronize_jf(); oracle.vpo_synchronize_jf(C.byref(a))
        assert np.array_equal(bits(f2), bits(W.fields))
        W.advance_b(0.5); oracle.vpo_advance_b(C.byref(a), 0.5)
        assert np.array_equal(bits(f2), bits(W.fields))
        W.advance_e(1.0); oracle.vpo_vacuum_advance_e(C.byref(a), 1.0)
        assert np.array_equal(bits(f2), bits(W.fields))
        W.advance_b(0.5); oracle.vpo_advance_b(C.byref(a), 0.5)
        assert np.array_equal(bits(f2), bits(W.fields))
    en = (C.c_double * 6)()
    oracle.vpo_vacuum_energy_f(C.byref(a), en)
    assert np.array_equal(np.array(en[:]), W.energy_f())
    W.clear_jf(); oracle.vpo_clear_jf(C.byref(a))
    assert np.array_equal(bits(f2), bits(W.fields))


@pytest.mark.parametrize("dims,fbc,material", [
    ((6, 5, 4), None, None),
    ((8, 8, 1), {0: -1, 3: -1}, None),                 # harris-like: pec x walls, degenerate z
    ((5, 1, 7), {2: -2, 5: -3}, None),                 # symmetric / pmc walls, degenerate y
    ((6, 5, 4), {0: -4, 3: -4, 2: -4, 5: -1}, None),   # absorbing walls on -x, +x, -z; pec on +z
    ((7, 4, 5), {0: -1, 3: -4, 1: -2, 4: -3}, DIELECTRIC),
])
def test_divergence_cleaning_bit_exact(ref_scalar, oracle, dims, fbc, material):
    """advance.cc:138-176: clear_rhof / synchronize_rho / compute_div_e_err / clean_div_e, compute_div_b_err /
    clean_div_b and synchronize_tang_e_norm_b, field array compared bit for bit after every call; the rms values are
    double sums whose grouping depends on the pipeline count, compared to 1e-13."""
    rng = np.random.default_rng(17)
    nx, ny, nz = dims
    W = R.RefWorld(ref_scalar, nx, ny, nz, fbc=fbc, material=material)
    f0 = rng.normal(0, 0.05, (W.nv, 20)).astype(np.float32)      # every slot, ghosts included, holds something
    f0[:, 16:] = 0
    W.fields[:] = f0
    f2 = f0.copy()
    a = W.field_args(f2)
    pa = C.byref(a)

    def same(what):
        assert np.array_equal(bits(f2), bits(W.fields)), what

    W.synchronize_rho(); oracle.vpo_synchronize_rho(pa); same("synchronize_rho")
    for rnd in range(3):
        W.compute_div_e_err(); oracle.vpo_vacuum_compute_div_e_err(pa); same("compute_div_e_err")
        np.testing.assert_allclose(oracle.vpo_compute_rms_div_e_err(pa), W.compute_rms_div_e_err(), rtol=1e-13)
        W.clean_div_e(); oracle.vpo_vacuum_clean_div_e(pa); same("clean_div_e")
    for rnd in range(3):
        W.compute_div_b_err(); oracle.vpo_compute_div_b_err(pa); same("compute_div_b_err")
        np.testing.assert_allclose(oracle.vpo_compute_rms_div_b_err(pa), W.compute_rms_div_b_err(), rtol=1e-13)
        W.clean_div_b(); oracle.vpo_clean_div_b(pa); same("clean_div_b")
    e_ref = W.synchronize_tang_e_norm_b(); e_orc = oracle.vpo_synchronize_tang_e_norm_b(pa); same("synchronize_tang_e_norm_b")
    np.testing.assert_allclose(e_orc, e_ref, rtol=1e-13)
    W.clear_rhof(); oracle.vpo_clear_rhof(pa); same("clear_rhof")
    W.compute_rhob(); oracle.vpo_vacuum_compute_rhob(pa); same("compute_rhob")
    W.compute_curl_b(); oracle.vpo_vacuum_compute_curl_b(pa); same("compute_curl_b")


def test_rho_p_and_rhob_bit_exact(ref_scalar, oracle):
    rng = np.random.default_rng(41)
    nx, ny, nz = 5, 4, 6
    W = make_world(ref_scalar, rng, nx, ny, nz)
    g = W.g.contents
    sp = W.new_species("r%d" % rng.integers(1 << 30), -1.0, 1.0, 8192, 16)
    parts = R.random_particles(rng, 6000, nx, ny, nz, w=0.3)
    sp.set_particles(parts)
    f2 = W.fields.copy()
    ref_scalar.accumulate_rho_p(W.fa, sp.sp)
    oracle.vpo_accumulate_rho_p(f2.ctypes.data, parts.ctypes.data, len(parts), -1.0, g.r8V, nx, ny, nz)
    assert np.array_equal(bits(f2), bits(W.fields))
    for k in range(200):                                   # includes wall voxels (doubled weights)
        ref_scalar.accumulate_rhob(W.fa.contents.f, parts[k:k + 1].ctypes.data, W.g, 1.5)
        oracle.vpo_accumulate_rhob(f2.ctypes.data, parts[k:k + 1].ctypes.data, 1.5, g.r8V, nx, ny, nz)
    assert np.array_equal(bits(f2), bits(W.fields))


@pytest.mark.parametrize("dims,fbc", [((6, 5, 4), None), ((7, 1, 5), {0: -1, 3: -4, 2: -2, 5: -3})])
def test_hydro_moments_bit_exact(ref_scalar, oracle, dims, fbc):
    """accumulate_hydro_p + synchronize_hydro_array: with a particle count that is a multiple of 16 the single
    pipeline sums every particle in array order, which is what the oracle does, so the comparison is bit for bit."""
    rng = np.random.default_rng(23)
    nx, ny, nz = dims
    W = R.RefWorld(ref_scalar, nx, ny, nz, fbc=fbc)
    W.fields[:] = R.random_fields(rng, W.nv)
    ref_scalar.load_interpolator_array(W.ia, W.fa)
    g = W.g.contents
    sp = W.new_species("h%d" % rng.integers(1 << 30), -1.0, 3.0, 8192, 16)
    parts = R.random_particles(rng, 4096, nx, ny, nz, uth=0.4, w=0.3)
    sp.set_particles(parts)
    ha = ref_scalar.new_hydro_array(W.g)
    ref_scalar.clear_hydro_array(ha)
    ref_scalar.accumulate_hydro_p(ha, sp.sp, W.ia)
    h_ref = np.ctypeslib.as_array(ha.contents.h, shape=(ha.contents.stride, 16))
    h2 = np.zeros((W.nv, 16), np.float32)
    f32 = np.float32
    qdt_2mc = f32(f32(f32(-1.0) * f32(g.dt)) / f32(f32(2) * f32(3.0) * f32(g.cvac)))
    oracle.vpo_accumulate_hydro_p(h2.ctypes.data, parts.ctypes.data, len(parts), W.interp.ctypes.data, W.isf,
                                  -1.0, 3.0, qdt_2mc, g.cvac, g.r8V, nx, ny, nz)
    ref_scalar.synchronize_hydro_array(ha)            # reduces the pipeline blocks, then walls and periodic folds
    a = W.field_args(W.fields)
    oracle.vpo_synchronize_hydro(h2.ctypes.data, C.byref(a))
    assert np.abs(h2[:, :14]).max() > 0
    assert np.array_equal(bits(h2[:, :14]), bits(h_ref[:W.nv, :14]))


@pytest.mark.parametrize("deck", ["accel", "cyclo", "inbndj", "interpe", "outbndj"])
def test_reference_build_passes_its_own_kat_decks(deck):
    """The reference as built by oracle/Makefile (MPI shim, plain g++) is only a valid yardstick if it still passes
    its own known-answer decks (test/integrated/legacy) — on the CPU, nothing of this repo on the path."""
    import subprocess, tempfile
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", f"{deck}.scalar")
    if not os.path.exists(path):
        pytest.skip("deck binaries not built (needs /root/reference at build time)")
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([path, "1", "1"], cwd=d, capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "pass" in out and "FAIL" not in out, out[-1500:]


def test_reference_pcomm_deck_passes_on_eight_shim_ranks():
    """test/integrated/legacy/pcomm.deck is the reference's own known-answer test of particle migration: eight MPI
    ranks (2 x 2 x 2), particles aimed across faces, edges and corners, exact voxels and offsets within 11 ulp on
    arrival.  It needs MPI; here the eight ranks are processes on oracle/mpi_shim's shared-memory transport
    (oracle/mpi_shim/shimrun).  Passing it pins that transport — the multi-rank oracle of the multi-GPU path."""
    import subprocess, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "oracle", "_ref", "pcomm.scalar")
    if not os.path.exists(path):
        pytest.skip("deck binaries not built (needs /root/reference at build time)")
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([os.path.join(root, "oracle", "mpi_shim", "shimrun"), "-n", "8", path, "1", "1"],
                           cwd=d, capture_output=True, text=True, timeout=600)
        logs = "".join(open(os.path.join(d, f)).read() for f in sorted(os.listdir(d)) if f.startswith("shimrun."))
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "pass" in out and "FAIL" not in out + logs, (out + logs)[-1500:]
    assert "8 (MPI) ranks" in out


def test_reference_build_passes_its_own_golden_energy_test():
    """test/unit/energy_comparison/3d_test against energies_gold.3d_test, reference alone on the CPU."""
    import shutil, subprocess, tempfile
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    exe, gold = os.path.join(ref, "3d_test.scalar"), os.path.join(ref, "energies_gold.3d_test")
    if not (os.path.exists(exe) and os.path.exists(gold)):
        pytest.skip("golden test binary not built (needs /root/reference at build time)")
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(gold, d)
        r = subprocess.run([exe, "--tpp", "1"], cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "All tests passed" in r.stdout + r.stderr, (r.stdout + r.stderr)[-1500:]


def _grid_heating(preload_env=None, timeout=1800):
    """Run test/unit/grid_heating (27 000 steps of a hot 2-D electron plasma) and the reference authors' checker."""
    import subprocess, sys, tempfile
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    exe, chk = os.path.join(ref, "gridHeatingTestElec.scalar"), os.path.join(ref, "grid_heating_check.py")
    if not (os.path.exists(exe) and os.path.exists(chk)):
        pytest.skip("grid heating test not built (needs /root/reference at build time)")
    env = dict(os.environ, **(preload_env or {}))
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([exe, "--tpp", "1"], cwd=d, env=env, capture_output=True, text=True, timeout=timeout)
        assert r.returncode == 0 and "1 passed" in r.stdout + r.stderr, (r.stdout + r.stderr)[-1500:]
        c = subprocess.run([sys.executable, chk, d], capture_output=True, text=True, timeout=300)
    return r.stdout + r.stderr, c.stdout + c.stderr


def test_reference_build_passes_its_own_grid_heating_test():
    """The heating rate of the reference as built here lies within the 5 standard deviations its authors allow."""
    _, verdict = _grid_heating()
    assert "Electron heating rate test PASS" in verdict, verdict[-1500:]

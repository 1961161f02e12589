"""GPU parity: the CUDA path (through the C-ABI, libvpic_b200.so) against the CPU oracle on identical seeded
inputs.  Integer/index work is bit-exact; particle state after one push is bit-exact (the kernels follow the scalar
reference operation for operation); sums whose order differs (atomics, parallel reductions) carry a stated tolerance.
"""
import ctypes as C
import numpy as np
import pytest

import refvpic as R
from vpic_b200 import abi, grid as G

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vpic_b200 import engine, lib
    lib.load()
    return engine


def make_grid(nx, ny, nz, pbc=None, fbc=None, frac=0.98):
    dt = G.courant_dt(1.0, 1.0, 1.0, nx, ny, nz, frac=frac)
    g = G.partition_periodic_box(0, 0, 0, nx, ny, nz, nx, ny, nz, 1, 1, 1, dt=dt)
    for f, c in (pbc or {}).items():
        g.set_pbc(f, c)
    for f, c in (fbc or {}).items():
        g.set_fbc(f, c)
    return g


def oracle_push(orc, g, parts, interp, q, m, max_nm, isf=20, asf=12):
    p2 = parts.copy()
    pm2 = np.zeros(max_nm, dtype=abi.mover_dtype)
    acc2 = np.zeros(((g.nv + 1) // 2 * 2, asf), dtype=np.float32)
    f32 = np.float32
    qf, mf, dt, cvac = f32(q), f32(m), f32(g.dt), f32(g.cvac)
    a = R.OraclePushArgs(p2.ctypes.data, len(p2), pm2.ctypes.data, max_nm, interp.ctypes.data, isf,
                         acc2.ctypes.data, asf, g.neighbor.ctypes.data, g.rangel, g.rangeh,
                         f32(f32(qf * dt) / f32(f32(f32(2) * mf) * cvac)),
                         f32(f32(cvac * dt) * f32(g.rdx)), f32(f32(cvac * dt) * f32(g.rdy)),
                         f32(f32(cvac * dt) * f32(g.rdz)), qf)
    ign = C.c_int32(0)
    nm = orc.vpo_advance_p(C.byref(a), C.byref(ign))
    return p2, pm2[:nm], acc2, ign.value


def accum_close(a_gpu, a_ref, rtol=2e-5):
    # fp32 sums of O(ppc) terms in a different order: compare against the magnitude of the entries
    scale = max(np.abs(a_ref).max(), 1e-30)
    err = np.abs(a_gpu - a_ref).max() / scale
    assert err < rtol, f"accumulator mismatch {err:.3e}"


PUSH_CASES = [
    ((6, 5, 4), 0.5, None, 20000, True),          # periodic, ~40% crossings, sorted input
    ((6, 5, 4), 0.5, None, 20000, False),         # same, unsorted input
    ((16, 1, 16), 0.3, None, 30011, True),        # 2-D deck shape, ragged np
    ((5, 4, 3), 0.6, {0: -1, 3: -1}, 9000, True),    # reflecting x walls
    ((5, 4, 3), 0.6, {2: -2, 5: -2}, 9000, True),    # absorbing z walls -> movers
    ((4, 4, 4), 0.1, None, 1, True),              # single particle
    ((4, 4, 4), 0.1, None, 0, True),              # empty species
]


@pytest.mark.parametrize("use_rule", [True, False])
@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("dims,uth,pbc,n,sort_first", PUSH_CASES)
def test_advance_p(eng, oracle, dims, uth, pbc, n, sort_first, variant, use_rule):
    rng = np.random.default_rng(17)
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz, pbc=pbc)
    fields = R.random_fields(rng, g.nv)
    interp = np.zeros((g.nv, 20), dtype=np.float32)
    oracle.vpo_load_interpolator(interp.ctypes.data, 20, fields.ctypes.data, nx, ny, nz)
    parts = R.random_particles(rng, n, nx, ny, nz, uth=uth, w=0.7)
    if sort_first and n:
        parts = parts[np.argsort(parts["i"], kind="stable")]
    max_nm = max(16, n)
    p_ref, pm_ref, acc_ref, _ = oracle_push(oracle, g, parts, interp, -1.0, 1.0, max_nm)

    dg = eng.DeviceGrid(g)
    dg.use_neighbor_rule = use_rule                      # closed-form neighbours vs the grid_t.neighbor table
    if use_rule:
        assert dg.neighbor_rule() is not None, "the closed form must verify on every grid the reference builds"
    ia, aa = eng.InterpolatorArray(dg), eng.AccumulatorArray(dg)
    ia.i.copy_(torch.from_numpy(interp))
    sp = eng.Species("electron", -1.0, 1.0, max(n, 1), max_nm, 20, 0, dg)
    sp.set_particles(parts)
    eng.clear_accumulator_array(aa)
    eng.advance_p(sp, aa, ia, variant=variant)
    eng.reduce_accumulator_array(aa)

    assert sp.nm == len(pm_ref) and sp.n_ignored == 0
    got = sp.particles_host()
    assert np.array_equal(bits(got), bits(p_ref)), "particle state must be bit-exact"
    assert np.array_equal(bits(sp.movers_host()), bits(pm_ref)), "movers must be bit-exact and ascending"
    accum_close(aa.a.cpu().numpy(), acc_ref)
    if n > 100:
        assert (got["i"] != parts["i"]).mean() > 0.05


@pytest.mark.parametrize("variant", [1, 2])
def test_advance_p_multi_span_and_grid_cap(eng, oracle, monkeypatch, variant):
    """The paths the benchmark runs at 134 M particles, forced at 300 k: spans of ONE row and a grid of one CTA per SM
    (args.debug_skip bits 16-23 and 24-31), so every warp takes several spans, prefetches across span boundaries and
    carries its mover queue from one span to the next."""
    monkeypatch.setenv("VPB_DEBUG_SKIP", str((1 << 16) | (1 << 24)))
    rng = np.random.default_rng(29)
    nx, ny, nz = 12, 10, 9
    g = make_grid(nx, ny, nz, pbc={2: -2, 5: -2})               # absorbing z walls: movers are emitted too
    fields = R.random_fields(rng, g.nv)
    interp = np.zeros((g.nv, 20), dtype=np.float32)
    oracle.vpo_load_interpolator(interp.ctypes.data, 20, fields.ctypes.data, nx, ny, nz)
    n = 300007
    parts = R.random_particles(rng, n, nx, ny, nz, uth=0.4, w=0.7)
    parts = parts[np.argsort(parts["i"], kind="stable")]
    p_ref, pm_ref, acc_ref, _ = oracle_push(oracle, g, parts, interp, -1.0, 1.0, n)
    dg = eng.DeviceGrid(g)
    ia, aa = eng.InterpolatorArray(dg), eng.AccumulatorArray(dg)
    ia.i.copy_(torch.from_numpy(interp))
    sp = eng.Species("electron", -1.0, 1.0, n, n, 20, 0, dg)
    sp.set_particles(parts)
    eng.clear_accumulator_array(aa)
    eng.advance_p(sp, aa, ia, variant=variant)
    assert sp.nm == len(pm_ref) > 0
    assert np.array_equal(bits(sp.particles_host()), bits(p_ref))
    assert np.array_equal(bits(sp.movers_host()), bits(pm_ref))
    accum_close(aa.a.cpu().numpy(), acc_ref)


def test_neighbor_rule_rejects_irregular_table(eng):
    """A neighbour table that is not the regular structure must fail verification (the kernels then use the table)."""
    g = make_grid(5, 4, 3)
    g.neighbor[G.voxel(2, 2, 2, 5, 4, 3), 3] = g.rangel + G.voxel(4, 4, 3, 5, 4, 3)     # a wormhole
    dg = eng.DeviceGrid(g)
    assert dg.neighbor_rule() is None


def test_advance_p_mover_overflow(eng, oracle):
    """More leavers than max_nm: extra movers are dropped, p.i stays a valid voxel (advance_p_pipeline.cc:223-236)."""
    rng = np.random.default_rng(2)
    nx, ny, nz = 4, 4, 4
    g = make_grid(nx, ny, nz, pbc={i: -2 for i in range(6)})
    interp = np.zeros((g.nv, 20), dtype=np.float32)
    parts = R.random_particles(rng, 5000, nx, ny, nz, uth=0.8)
    dg = eng.DeviceGrid(g)
    ia, aa = eng.InterpolatorArray(dg), eng.AccumulatorArray(dg)
    sp = eng.Species("e", -1.0, 1.0, 5000, 8, 20, 0, dg)
    sp.set_particles(parts)
    with pytest.warns(UserWarning):
        eng.advance_p(sp, aa, ia)
    assert sp.nm == 8 and sp.n_ignored > 0
    got = sp.particles_host()
    mv = sp.movers_host()
    waiting = np.zeros(len(got), bool); waiting[mv["i"]] = True
    assert np.all(got["i"][~waiting] < g.nv) and np.all(got["i"][~waiting] >= 0)
    assert np.all(got["i"][waiting] >= 8)


@pytest.mark.parametrize("dims,n", [((7, 6, 5), 17777), ((32, 32, 32), 300000), ((64, 64, 1), 262144),
                                    ((3, 3, 3), 1), ((3, 3, 3), 0), ((130, 70, 40), 50000)])
def test_sort_p(eng, oracle, dims, n):
    rng = np.random.default_rng(23)
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz)
    parts = R.random_particles(rng, n, nx, ny, nz)
    parts["w"] = np.arange(n, dtype=np.float32)           # tag to expose any instability
    p_ref, aux = parts.copy(), np.zeros_like(parts)
    part_ref = np.zeros(g.nv + 1, dtype=np.int32)
    oracle.vpo_sort_p(p_ref.ctypes.data, n, aux.ctypes.data, part_ref.ctypes.data, nx, ny, nz)
    dg = eng.DeviceGrid(g)
    sp = eng.Species("e", -1.0, 1.0, max(n, 1), 16, 20, 0, dg)
    sp.set_particles(parts)
    eng.sort_p(sp)
    assert np.array_equal(bits(sp.particles_host()), bits(p_ref)), "sort order must be bit-exact (stable)"
    assert np.array_equal(sp.partition.cpu().numpy()[:g.nv], part_ref[:g.nv])
    assert sp.last_sorted == g.step
    # idempotence: sorting sorted data changes nothing
    eng.sort_p(sp)
    assert np.array_equal(bits(sp.particles_host()), bits(p_ref))


def test_sort_p_at_scale(eng, oracle):
    """12 M particles: every CTA of the scatter kernel walks several sub-tiles (re-zeroing its touched-digit bitmap in
    between) — the path the benchmark's 134 M-particle sorts take.  Order and partition[] bit-exact vs the oracle."""
    nx, ny, nz = 64, 48, 40
    n = 12_000_003
    g = make_grid(nx, ny, nz)
    gen = np.random.default_rng(8)
    parts = np.zeros(n, dtype=abi.particle_dtype)
    parts["i"] = (gen.integers(1, nx + 1, n) + (nx + 2) * (gen.integers(1, ny + 1, n) + (ny + 2) * gen.integers(1, nz + 1, n))).astype(np.int32)
    parts["w"] = np.arange(n, dtype=np.float32)                # tags (exact in fp32 up to 2^24) expose instability
    parts["ux"] = gen.random(n, dtype=np.float32)
    p_ref, aux = parts.copy(), np.zeros_like(parts)
    part_ref = np.zeros(g.nv + 1, dtype=np.int32)
    oracle.vpo_sort_p(p_ref.ctypes.data, n, aux.ctypes.data, part_ref.ctypes.data, nx, ny, nz)
    del aux
    dg = eng.DeviceGrid(g)
    sp = eng.Species("e", -1.0, 1.0, n, 16, 20, 0, dg)
    sp.set_particles(parts)
    eng.sort_p(sp)
    assert np.array_equal(bits(sp.particles_host()), bits(p_ref)), "sort order must be bit-exact (stable) at scale"
    assert np.array_equal(sp.partition.cpu().numpy()[:g.nv], part_ref[:g.nv])


@pytest.mark.parametrize("dims,n", [((7, 6, 5), 17777), ((32, 32, 32), 300000), ((64, 64, 1), 262144), ((3, 3, 3), 1),
                                    ((3, 3, 3), 0), ((130, 70, 40), 50000), ((300, 300, 64), 200000), ((40, 30, 20), 3_000_001)])
def test_sort_p_deferred_settles_to_the_reference_order(eng, oracle, dims, n):
    """sort_p(defer=True) sorts (voxel, index) pairs and leaves the particles in place; reading the array (or any other
    operator) applies the order.  Order and partition[] bit-exact vs the oracle, for one-, two- and three-pass digit plans
    (300x300x64 cells need 23 key bits) and for CTAs that walk several sub-tiles (3 M particles)."""
    rng = np.random.default_rng(29)
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz)
    parts = R.random_particles(rng, n, nx, ny, nz)
    parts["w"] = np.arange(n, dtype=np.float32)           # tag to expose any instability
    p_ref, aux = parts.copy(), np.zeros_like(parts)
    part_ref = np.zeros(g.nv + 1, dtype=np.int32)
    oracle.vpo_sort_p(p_ref.ctypes.data, n, aux.ctypes.data, part_ref.ctypes.data, nx, ny, nz)
    dg = eng.DeviceGrid(g)
    sp = eng.Species("e", -1.0, 1.0, max(n, 1), 16, 20, 0, dg)
    sp.set_particles(parts)
    eng.sort_p(sp, defer=True)
    assert sp._perm_pending == (n > 1)
    assert np.array_equal(sp.partition.cpu().numpy()[:g.nv], part_ref[:g.nv])
    if n > 1:
        assert np.array_equal(bits(sp._p[:n].cpu().numpy().reshape(-1).view(abi.particle_dtype)), bits(parts)), "not moved yet"
    assert np.array_equal(bits(sp.particles_host()), bits(p_ref)), "settled order must be bit-exact (stable)"
    assert not sp._perm_pending
    eng.sort_p(sp, defer=True)                            # sorted input: the identity order
    assert np.array_equal(bits(sp.particles_host()), bits(p_ref))


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_sort_p_deferred_fused_into_advance_p(eng, oracle, case):
    """The advance_p that follows a deferred sort_p loads p[perm[k]] and stores position k of the other buffer: particle
    bytes, movers and partition[] identical to sort_p followed by advance_p, accumulators to summation-order tolerance.
    Cases: periodic box; absorbing walls (movers leave the domain); few particles per row of a big grid; the span and
    grid overrides do not apply here (debug switches are off in this variant), so a 1.2 M-particle case covers warps
    that take several spans."""
    rng = np.random.default_rng(31 + case)
    dims, n, pbc, uth = [((12, 10, 8), 40011, None, 0.3), ((9, 7, 6), 20000, {i: -2 for i in range(6)}, 0.6),
                         ((48, 40, 36), 30001, None, 0.25), ((24, 20, 16), 1_200_007, None, 0.2)][case]
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz, pbc=pbc) if pbc else make_grid(nx, ny, nz)
    fields = R.random_fields(rng, g.nv)
    parts = R.random_particles(rng, n, nx, ny, nz, uth=uth)
    dg = eng.DeviceGrid(g)
    fa, ia = eng.FieldArray(dg), eng.InterpolatorArray(dg)
    fa.f.copy_(torch.from_numpy(fields))
    eng.load_interpolator_array(ia, fa)
    out = []
    for defer in (False, True):
        aa = eng.AccumulatorArray(dg)
        sp = eng.Species("e", -1.0, 1.0, n, n, 20, 0, dg)
        sp.set_particles(parts)
        eng.clear_accumulator_array(aa)
        eng.sort_p(sp, defer=defer)
        eng.advance_p(sp, aa, ia)
        assert not sp._perm_pending
        out.append((sp.particles_host().copy(), sp.movers_host().copy(), sp.nm, sp.partition.cpu().numpy().copy(),
                    aa.a.cpu().numpy().copy()))
    (p0, m0, nm0, part0, a0), (p1, m1, nm1, part1, a1) = out
    assert nm0 == nm1 and (pbc is None or nm0 > 0)
    assert np.array_equal(bits(p0), bits(p1)), "particles after the fused sort+push must be bit-identical"
    assert np.array_equal(bits(m0), bits(m1))
    assert np.array_equal(part0, part1)
    scale = max(np.abs(a0).max(), 1e-30)
    assert np.abs(a0 - a1).max() <= 2e-5 * scale


@pytest.mark.parametrize("case", [0, 1])
def test_sort_p_deferred_takes_its_keys_from_the_previous_push(eng, oracle, case):
    """advance_p(emit_keys=True) leaves every particle's voxel in a compact array; the deferred sort_p that follows sorts
    those and never reads the particles.  Two steps (push, sort, push) with and without the short cut: identical bytes.
    Case 1 has absorbing walls: movers leave the domain, so the keys must be ignored (sp.nm != 0)."""
    rng = np.random.default_rng(41 + case)
    dims, n, pbc, uth = [((14, 12, 10), 150011, None, 0.35), ((9, 7, 6), 20000, {i: -2 for i in range(6)}, 0.6)][case]
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz, pbc=pbc) if pbc else make_grid(nx, ny, nz)
    fields = R.random_fields(rng, g.nv)
    parts = R.random_particles(rng, n, nx, ny, nz, uth=uth)
    dg = eng.DeviceGrid(g)
    fa, ia = eng.FieldArray(dg), eng.InterpolatorArray(dg)
    fa.f.copy_(torch.from_numpy(fields))
    eng.load_interpolator_array(ia, fa)
    out = []
    for fast in (False, True):
        aa = eng.AccumulatorArray(dg)
        sp = eng.Species("e", -1.0, 1.0, n, n, 20, 0, dg)
        sp.set_particles(parts)
        eng.clear_accumulator_array(aa)
        eng.advance_p(sp, aa, ia, emit_keys=fast)
        used_keys = fast and sp._keys_valid and sp.nm == 0
        if sp.nm:                                            # what boundary_p does on one rank: absorbed particles go
            eng.boundary_pack(sp, [-1] * 6, fa)
        eng.sort_p(sp, defer=fast)
        eng.advance_p(sp, aa, ia)
        out.append((sp.particles_host().copy(), sp.np, sp.partition.cpu().numpy().copy(), used_keys))
    (p0, np0, part0, _), (p1, np1, part1, used) = out
    assert used == (case == 0)
    assert np0 == np1 and np.array_equal(bits(p0), bits(p1)) and np.array_equal(part0, part1)


def test_sort_movers(eng):
    from vpic_b200 import lib
    L = lib.load()
    rng = np.random.default_rng(4)
    n = 70001
    mv = np.zeros(n, dtype=abi.mover_dtype)
    mv["i"] = rng.permutation(5_000_000)[:n].astype(np.int32)
    mv["dispx"] = mv["i"].astype(np.float32)
    t = torch.from_numpy(mv.view(np.float32).reshape(-1, 4)).cuda()
    need = L.vpb_sort_movers_scratch_bytes(n)
    scratch = torch.empty(need, dtype=torch.uint8, device="cuda")
    lib.check(L.vpb_sort_movers(t.data_ptr(), n, scratch.data_ptr(), need, None))
    out = t.cpu().numpy().reshape(-1).view(abi.mover_dtype)
    assert np.array_equal(out["i"], np.sort(mv["i"]))
    assert np.array_equal(out["dispx"], out["i"].astype(np.float32))


@pytest.mark.parametrize("dims", [(6, 4, 5), (64, 64, 1), (33, 17, 9)])
def test_interpolator_accumulator_glue(eng, oracle, dims):
    rng = np.random.default_rng(9)
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz)
    dg = eng.DeviceGrid(g)
    fields = R.random_fields(rng, g.nv)
    fa, ia, aa = eng.FieldArray(dg), eng.InterpolatorArray(dg), eng.AccumulatorArray(dg)
    fa.f.copy_(torch.from_numpy(fields))
    eng.load_interpolator_array(ia, fa)
    i_ref = np.zeros((g.nv, 20), dtype=np.float32)
    oracle.vpo_load_interpolator(i_ref.ctypes.data, 20, fields.ctypes.data, nx, ny, nz)
    assert np.array_equal(bits(ia.i.cpu().numpy()), bits(i_ref))

    acc = np.zeros((aa.stride, 12), dtype=np.float32)
    x, y, z = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    v = abi.voxel(x, y, z, nx, ny, nz).ravel()
    acc[v] = rng.normal(0, 1, (len(v), 12)).astype(np.float32)
    aa.a.copy_(torch.from_numpy(acc))
    f_ref = fields.copy()
    oracle.vpo_unload_accumulator(f_ref.ctypes.data, acc.ctypes.data, 12, nx, ny, nz, g.rdx, g.rdy, g.rdz, g.dt)
    eng.unload_accumulator_array(fa, aa)
    assert np.array_equal(bits(fa.f.cpu().numpy()), bits(f_ref))

    a_ref = acc.copy()
    a_ref[0] = 5.0                                             # ghost voxel below the cleared window stays untouched
    aa.a.copy_(torch.from_numpy(a_ref))
    oracle.vpo_clear_accumulator(a_ref.ctypes.data, 12, nx, ny, nz)
    eng.clear_accumulator_array(aa)
    assert np.array_equal(bits(aa.a.cpu().numpy()), bits(a_ref))


def test_energy_center_uncenter(eng, oracle):
    rng = np.random.default_rng(31)
    nx, ny, nz = 8, 6, 5
    g = make_grid(nx, ny, nz)
    dg = eng.DeviceGrid(g)
    fields = R.random_fields(rng, g.nv)
    interp = np.zeros((g.nv, 20), dtype=np.float32)
    oracle.vpo_load_interpolator(interp.ctypes.data, 20, fields.ctypes.data, nx, ny, nz)
    ia = eng.InterpolatorArray(dg)
    ia.i.copy_(torch.from_numpy(interp))
    parts = R.random_particles(rng, 40001, nx, ny, nz, uth=0.4, w=0.37)
    sp = eng.Species("ion", 1.0, 25.0, len(parts), 16, 20, 0, dg)
    sp.set_particles(parts)
    e_ref = oracle.vpo_energy_p(parts.ctypes.data, len(parts), interp.ctypes.data, 20, 1.0, 25.0, g.dt, g.cvac)
    e_gpu = eng.energy_p(sp, ia)
    assert abs(e_gpu - e_ref) <= 1e-12 * abs(e_ref)            # same fp32 terms, double sums in another order
    qdt_2mc = sp.push_constants()[0]
    p_ref = parts.copy()
    oracle.vpo_uncenter_p(p_ref.ctypes.data, len(p_ref), interp.ctypes.data, 20, qdt_2mc)
    eng.uncenter_p(sp, ia)
    assert np.array_equal(bits(sp.particles_host()), bits(p_ref))
    oracle.vpo_center_p(p_ref.ctypes.data, len(p_ref), interp.ctypes.data, 20, qdt_2mc)
    eng.center_p(sp, ia)
    assert np.array_equal(bits(sp.particles_host()), bits(p_ref))


# material_coefficient_t of one anisotropic conducting material (decay/drive x,y,z; rmu x,y,z; nonconductive; eps x,y,z)
MATERIAL = [0.93, 0.64, 0.88, 0.47, 0.97, 0.39, 0.83, 0.91, 0.77, 0.0, 1.5, 2.0, 2.5]


@pytest.mark.parametrize("dims,fbc,damp,material", [
    ((6, 5, 4), None, 0.0, None),
    ((6, 5, 4), None, 0.01, None),
    ((64, 64, 1), {0: -1, 3: -1}, 0.0, None),          # harris: pec x walls, one cell in z
    ((5, 1, 7), {2: -2, 5: -3}, 0.0, None),            # symmetric / pmc, one cell in y
    ((40, 24, 16), None, 0.0, None),
    ((6, 5, 4), {0: -4, 3: -4, 2: -4, 5: -1}, 0.0, None),   # absorbing (Higdon) walls
    ((96, 1, 40), {0: -4, 3: -4}, 0.01, None),          # lpi-like 2-D box with absorbing x walls
    ((33, 9, 12), {0: -1, 3: -4}, 0.01, MATERIAL),      # one non-vacuum material filling space (sfa.cc:202-211)
])
def test_field_advance(eng, oracle, dims, fbc, damp, material):
    rng = np.random.default_rng(3)
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz, fbc=fbc)
    dg = eng.DeviceGrid(g)
    f0 = R.random_fields(rng, g.nv)
    f0[:, 8:11] = rng.normal(0, 0.01, (g.nv, 3))
    f0[:, 12:15] = rng.normal(0, 0.02, (g.nv, 3))
    fa = eng.FieldArray(dg, damp=damp, material=material)
    fa.f.copy_(torch.from_numpy(f0))
    f_ref = f0.copy()
    a = R.OracleFieldArgs()
    if material is not None:
        a.has_material = 1
        for i, v in enumerate(material):
            a.material[i] = v
    a.f = f_ref.ctypes.data
    a.nx, a.ny, a.nz = nx, ny, nz
    a.dt, a.cvac, a.eps0, a.damp = g.dt, g.cvac, g.eps0, damp
    a.dx, a.dy, a.dz, a.dV = g.dx, g.dy, g.dz, g.dV
    a.rdx, a.rdy, a.rdz = g.rdx, g.rdy, g.rdz
    for i, (fi, fj, fk) in enumerate(G.FACES):
        a.bc6[i] = g.bc[G.boundary_index(fi, fj, fk)]
    for _ in range(3):
        fa.synchronize_jf(); oracle.vpo_synchronize_jf(C.byref(a))
        assert np.array_equal(bits(fa.f.cpu().numpy()), bits(f_ref)), "synchronize_jf"
        fa.advance_b(0.5); oracle.vpo_advance_b(C.byref(a), 0.5)
        assert np.array_equal(bits(fa.f.cpu().numpy()), bits(f_ref)), "advance_b"
        fa.advance_e(1.0); oracle.vpo_vacuum_advance_e(C.byref(a), 1.0)
        assert np.array_equal(bits(fa.f.cpu().numpy()), bits(f_ref)), "advance_e"
        fa.advance_b(0.5); oracle.vpo_advance_b(C.byref(a), 0.5)
        assert np.array_equal(bits(fa.f.cpu().numpy()), bits(f_ref)), "advance_b"
    en = (C.c_double * 6)()
    oracle.vpo_vacuum_energy_f(C.byref(a), en)
    np.testing.assert_allclose(fa.energy_f(), np.array(en[:]), rtol=1e-12)
    fa.clear_jf(); oracle.vpo_clear_jf(C.byref(a))
    assert np.array_equal(bits(fa.f.cpu().numpy()), bits(f_ref))


@pytest.mark.parametrize("dims,fbc,material", [
    ((6, 5, 4), None, None),
    ((64, 64, 1), {0: -1, 3: -1}, None),               # harris: pec x walls, one cell in z
    ((5, 1, 7), {2: -2, 5: -3}, None),                 # symmetric / pmc walls, one cell in y
    ((40, 24, 16), None, None),
    ((6, 5, 4), {0: -4, 3: -4, 2: -4, 5: -1}, None),   # absorbing walls
    ((33, 9, 12), {0: -1, 3: -4, 1: -2, 4: -3}, MATERIAL),
])
def test_divergence_cleaning(eng, oracle, dims, fbc, material):
    """advance.cc:138-176 on the device against the oracle (itself pinned bit-exact against the reference): the field
    array bit for bit after every call, the three double-precision reductions to 1e-12."""
    rng = np.random.default_rng(17)
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz, fbc=fbc)
    dg = eng.DeviceGrid(g)
    f0 = rng.normal(0, 0.05, (g.nv, 20)).astype(np.float32)
    f0[:, 16:] = 0
    fa = eng.FieldArray(dg, material=material)
    fa.f.copy_(torch.from_numpy(f0))
    f_ref = f0.copy()
    a = R.OracleFieldArgs()
    a.f = f_ref.ctypes.data
    a.nx, a.ny, a.nz = nx, ny, nz
    a.dt, a.cvac, a.eps0, a.damp = g.dt, g.cvac, g.eps0, 0.0
    a.dx, a.dy, a.dz, a.dV = g.dx, g.dy, g.dz, g.dV
    a.rdx, a.rdy, a.rdz = g.rdx, g.rdy, g.rdz
    for i, (fi, fj, fk) in enumerate(G.FACES):
        a.bc6[i] = g.bc[G.boundary_index(fi, fj, fk)]
    if material is not None:
        a.has_material = 1
        for i, v in enumerate(material):
            a.material[i] = v
    pa = C.byref(a)

    def same(what):
        assert np.array_equal(bits(fa.f.cpu().numpy()), bits(f_ref)), what

    fa.synchronize_rho(); oracle.vpo_synchronize_rho(pa); same("synchronize_rho")
    for _ in range(3):
        fa.compute_div_e_err(); oracle.vpo_vacuum_compute_div_e_err(pa); same("compute_div_e_err")
        np.testing.assert_allclose(fa.compute_rms_div_e_err(), oracle.vpo_compute_rms_div_e_err(pa), rtol=1e-12)
        fa.clean_div_e(); oracle.vpo_vacuum_clean_div_e(pa); same("clean_div_e")
    for _ in range(3):
        fa.compute_div_b_err(); oracle.vpo_compute_div_b_err(pa); same("compute_div_b_err")
        np.testing.assert_allclose(fa.compute_rms_div_b_err(), oracle.vpo_compute_rms_div_b_err(pa), rtol=1e-12)
        fa.clean_div_b(); oracle.vpo_clean_div_b(pa); same("clean_div_b")
    err = fa.synchronize_tang_e_norm_b(); err_ref = oracle.vpo_synchronize_tang_e_norm_b(pa); same("synchronize_tang_e_norm_b")
    np.testing.assert_allclose(err, err_ref, rtol=1e-12, atol=1e-300)
    fa.clear_rhof(); oracle.vpo_clear_rhof(pa); same("clear_rhof")
    fa.compute_rhob(); oracle.vpo_vacuum_compute_rhob(pa); same("compute_rhob")
    fa.compute_curl_b(); oracle.vpo_vacuum_compute_curl_b(pa); same("compute_curl_b")


@pytest.mark.parametrize("dims,fbc,n,sort_first", [((6, 5, 4), None, 20011, True), ((6, 5, 4), None, 20011, False),
                                                   ((33, 1, 9), {0: -1, 3: -4, 2: -2, 5: -3}, 150001, True)])
def test_hydro_moments(eng, oracle, dims, fbc, n, sort_first):
    """accumulate_hydro_p + synchronize_hydro_array against the oracle (pinned bit-exact against the reference): the 14
    moments per node agree to fp32 summation-order tolerance; sorted input takes the warp-reduced path, unsorted the
    per-lane REDs."""
    rng = np.random.default_rng(29)
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz, fbc=fbc)
    dg = eng.DeviceGrid(g)
    fields = R.random_fields(rng, g.nv)
    parts = R.random_particles(rng, n, nx, ny, nz, uth=0.4, w=0.3)
    if sort_first:
        parts = parts[np.argsort(parts["i"], kind="stable")]
    fa, ia = eng.FieldArray(dg), eng.InterpolatorArray(dg)
    fa.f.copy_(torch.from_numpy(fields))
    eng.load_interpolator_array(ia, fa)
    sp = eng.Species("e", -1.0, 3.0, n, 16, 20, 0, dg)
    sp.set_particles(parts)
    ha = eng.HydroArray(dg)
    ha.h.fill_(7.0)                                              # clear must really clear
    ha.clear()
    eng.accumulate_hydro_p(ha, sp, ia)
    ha.synchronize(fa)
    interp = ia.i.cpu().numpy()
    h_ref = np.zeros((g.nv, 16), np.float32)
    f32 = np.float32
    qdt_2mc = f32(f32(f32(-1.0) * f32(g.dt)) / f32(f32(2) * f32(3.0) * f32(g.cvac)))
    oracle.vpo_accumulate_hydro_p(h_ref.ctypes.data, parts.ctypes.data, n, interp.ctypes.data, interp.shape[1],
                                  -1.0, 3.0, qdt_2mc, g.cvac, g.r8V, nx, ny, nz)
    a = R.OracleFieldArgs()
    a.f = h_ref.ctypes.data
    a.nx, a.ny, a.nz = nx, ny, nz
    a.dx, a.dy, a.dz = g.dx, g.dy, g.dz
    for i, (fi, fj, fk) in enumerate(G.FACES):
        a.bc6[i] = g.bc[G.boundary_index(fi, fj, fk)]
    oracle.vpo_synchronize_hydro(h_ref.ctypes.data, C.byref(a))
    got = ha.h.cpu().numpy()
    for col in range(14):
        scale = np.abs(h_ref[:, col]).max()
        assert scale > 0
        assert np.abs(got[:, col] - h_ref[:, col]).max() <= 3e-5 * scale, col
    assert np.array_equal(got[:, 14:], np.zeros_like(got[:, 14:]))


def test_reference_scalar_agrees_when_present(eng, oracle):
    """If the prebuilt unmodified reference travelled with the repo, check the CUDA push against it directly."""
    if not R.have_ref("scalar"):
        pytest.skip("oracle/_ref not present")
    lib = R.load_ref("scalar", tpp=1)
    rng = np.random.default_rng(77)
    nx, ny, nz = 6, 5, 4
    W = R.RefWorld(lib, nx, ny, nz)
    W.fields[:] = R.random_fields(rng, W.nv)
    lib.load_interpolator_array(W.ia, W.fa)
    rs = W.new_species("e_gpu_ref", -1.0, 1.0, 8192, 8192)
    parts = R.random_particles(rng, 8000, nx, ny, nz, uth=0.4)
    rs.set_particles(parts)
    lib.clear_accumulator_array(W.aa); lib.advance_p(rs.sp, W.aa, W.ia); lib.reduce_accumulator_array(W.aa)
    gc = W.g.contents
    g = G.partition_periodic_box(0, 0, 0, nx, ny, nz, nx, ny, nz, 1, 1, 1, dt=gc.dt)
    assert np.array_equal(g.neighbor, W.neighbor)
    dg = eng.DeviceGrid(g)
    ia, aa = eng.InterpolatorArray(dg), eng.AccumulatorArray(dg)
    ia.i.copy_(torch.from_numpy(W.interp.copy()))
    sp = eng.Species("e", -1.0, 1.0, 8192, 8192, 20, 0, dg)
    sp.set_particles(parts)
    eng.clear_accumulator_array(aa); eng.advance_p(sp, aa, ia)
    assert np.array_equal(bits(sp.particles_host()), bits(rs.p[:8000]))
    accum_close(aa.a.cpu().numpy(), W.accum[0])


@pytest.mark.parametrize("extra", [[], ["--axis", "0", "--harris"], ["--clean"], ["--axis", "0", "--harris", "--clean"]])
def test_multi_gpu_slab(eng, extra):
    """Slab-decomposed run over NCCL vs the undecomposed run (tests/multi_gpu_check.py); needs >= 2 GPUs.
    Second case: x slabs with conducting, particle-reflecting z walls and a sheared B field (C4-like)."""
    import os, subprocess, sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "multi_gpu_check.py")] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("mode", ["coherent", "resident", "auto"])
def test_dropin_symbols_on_host_structs(eng, oracle, mode):
    """The reference-named entry points (advance_p(species_t*, ...), sort_p, ...) on HOST structs: chunked, pipelined
    copies in coherent mode (chunk forced small so several chunks are in flight), explicit syncs in resident mode,
    and no syncs at all in auto mode — there the host reads below fault the device-owned pages back in."""
    import subprocess, sys, os, json, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent(f"""
        import sys, ctypes as C, numpy as np
        sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})
        import bench, refvpic as R
        from vpic_b200 import lib, grid as G, abi
        L = lib.load(); orc = R.load_oracle()
        nx, ny, nz, n = 7, 6, 5, 23003
        g = G.partition_periodic_box(0,0,0,nx,ny,nz,nx,ny,nz,1,1,1, dt=G.courant_dt(1,1,1,nx,ny,nz,frac=0.98))
        g.set_pbc(2, -2); g.set_pbc(5, -2)                      # absorbing z walls: movers come back to the host
        mode = {mode!r}
        H = bench.HostWorld(L, g, pinned='register' if mode == 'auto' else True)
        L.vpic_b200_set_lazy_min.argtypes = [C.c_size_t]; L.vpic_b200_set_lazy_min.restype = None
        L.vpic_b200_set_lazy_min(4096)
        L.vpic_b200_set_mode(dict(coherent=0, resident=1, auto=2)[mode])
        rng = np.random.default_rng(8)
        H.fields[:] = R.random_fields(rng, g.nv)
        sp = H.new_species('e', -1.0, 1.0, n, n, 20)
        parts = R.random_particles(rng, n, nx, ny, nz, uth=0.4, w=0.5)
        sp.p[:n] = parts.view(np.float32).reshape(-1, 8); sp.c.np = n
        H.load_interpolator()
        L.sort_p(C.byref(sp.c))
        L.clear_accumulator_array(C.byref(H.aa))
        L.advance_p(C.byref(sp.c), C.byref(H.aa), C.byref(H.ia))
        L.reduce_accumulator_array(C.byref(H.aa))
        L.unload_accumulator_array(C.byref(H.fa), C.byref(H.aa))
        for fn in ('vpic_b200_sync_to_host',):
            getattr(L, fn).argtypes = [C.c_void_p]; getattr(L, fn).restype = None
        if mode != 'auto':
            L.vpic_b200_sync_to_host(None)
        # oracle on the same inputs
        f0 = H.fields.copy(); f0[:, 12:15] = 0
        interp = np.zeros((g.nv, 20), np.float32)
        fld_in = R.random_fields(np.random.default_rng(8), g.nv)
        orc.vpo_load_interpolator(interp.ctypes.data, 20, fld_in.ctypes.data, nx, ny, nz)
        p2, aux = parts.copy(), np.zeros_like(parts); part = np.zeros(g.nv + 1, np.int32)
        orc.vpo_sort_p(p2.ctypes.data, n, aux.ctypes.data, part.ctypes.data, nx, ny, nz)
        pm2 = np.zeros(n, dtype=abi.mover_dtype); acc2 = np.zeros(((g.nv + 1)//2*2, 12), np.float32)
        f32 = np.float32; dt = f32(g.dt)
        a = R.OraclePushArgs(p2.ctypes.data, n, pm2.ctypes.data, n, interp.ctypes.data, 20, acc2.ctypes.data, 12,
                             g.neighbor.ctypes.data, g.rangel, g.rangeh, f32(f32(f32(-1)*dt)/f32(2)), dt, dt, dt, f32(-1))
        nm = orc.vpo_advance_p(C.byref(a), None)
        got = sp.p[:n].reshape(-1).view(abi.particle_dtype)
        ok = dict(interp=bool(np.array_equal(H.interp.view(np.uint32), interp.view(np.uint32))),
                  nm=bool(nm == sp.c.nm and nm > 0),
                  particles=bool(np.array_equal(got.view(np.uint8), p2.view(np.uint8))),
                  movers=bool(np.array_equal(sp.pm[:nm].reshape(-1).view(np.uint8), pm2[:nm].view(np.uint8))),
                  partition=bool(np.array_equal(sp.partition[:g.nv], part[:g.nv])),
                  accum=float(np.abs(H.accum - acc2).max() / np.abs(acc2).max()),
                  last_sorted=int(sp.c.last_sorted))
        fld_ref = fld_in.copy()
        orc.vpo_unload_accumulator(fld_ref.ctypes.data, H.accum.ctypes.data, 12, nx, ny, nz, g.rdx, g.rdy, g.rdz, g.dt)
        ok['jf'] = bool(np.array_equal(fld_ref.view(np.uint32), H.fields.view(np.uint32)))
        st = (C.c_uint64 * 4)(); L.vpic_b200_lazy_stats(st)
        ok['lazy'] = [int(x) for x in st]
        L.vpic_b200_set_mode(0)                                  # hands every array back before the buffers are freed
        import json; print('RESULT ' + json.dumps(ok))
    """)
    env = dict(os.environ, VPIC_B200_CHUNK="4096", VPIC_B200_LAZY_CHUNK="8192")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][0][7:])
    assert res["interp"] and res["nm"] and res["particles"] and res["movers"] and res["partition"] and res["jf"], res
    assert res["accum"] < 2e-5 and res["last_sorted"] == 0, res
    if mode == "auto":
        faults, fault_bytes, remaps, regions = res["lazy"]
        assert faults > 0 and fault_bytes > 0 and regions >= 4 and remaps == 0, res


_DEFER_SCRIPT = """
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import bench, refvpic as R
from vpic_b200 import lib, grid as G, abi
mode, scenario = sys.argv[1], sys.argv[2]
L = lib.load(); orc = R.load_oracle()
nx, ny, nz, n = 9, 8, 7, 70001
g = G.partition_periodic_box(0,0,0,nx,ny,nz,nx,ny,nz,1,1,1, dt=G.courant_dt(1,1,1,nx,ny,nz,frac=0.98))
H = bench.HostWorld(L, g, pinned='register')
L.vpic_b200_set_lazy_min.argtypes = [C.c_size_t]; L.vpic_b200_set_lazy_min.restype = None
L.vpic_b200_set_lazy_min(4096)
L.vpic_b200_sync_to_host.argtypes = [C.c_void_p]; L.vpic_b200_sync_to_host.restype = None
L.vpic_b200_set_mode(dict(resident=1, auto=2)[mode])
rng = np.random.default_rng(33)
fld = R.random_fields(rng, g.nv) * 0.05
H.fields[:] = fld
sp = H.new_species('e', -1.0, 1.0, n + 64, n, 3)
parts = R.random_particles(rng, n, nx, ny, nz, uth=0.3, w=0.5)
sp.p[:n] = parts.view(np.float32).reshape(-1, 8); sp.c.np = n
H.load_interpolator()
# oracle: sort, the same host edit, push
interp = np.zeros((g.nv, 20), np.float32)
orc.vpo_load_interpolator(interp.ctypes.data, 20, fld.ctypes.data, nx, ny, nz)
p2, aux = parts.copy(), np.zeros_like(parts); part = np.zeros(g.nv + 1, np.int32)
orc.vpo_sort_p(p2.ctypes.data, n, aux.ctypes.data, part.ctypes.data, nx, ny, nz)
sorted_ref = p2.copy()
L.sort_p(C.byref(sp.c))
seen = {{}}
edit = None
if scenario == 'edge_write':          # the first particles live in the unprotected head of the array (never faults)
    edit = [0, 3]
elif scenario == 'interior_read':     # a host read in the middle faults a chunk back: the order is applied first
    if mode == 'resident': L.vpic_b200_sync_to_host(sp.p.ctypes.data)
    k = n // 2
    seen['mid'] = bool(np.array_equal(sp.p[k].view(np.uint32), sorted_ref[k:k+1].view(np.uint32).reshape(-1)))
    edit = [k]
elif scenario == 'edge_read':
    if mode == 'resident': L.vpic_b200_sync_to_host(sp.p.ctypes.data)
    seen['first'] = bool(np.array_equal(sp.p[0].view(np.uint32), sorted_ref[0:1].view(np.uint32).reshape(-1)))
    seen['fourth'] = bool(np.array_equal(sp.p[3].view(np.uint32), sorted_ref[3:4].view(np.uint32).reshape(-1)))
if edit and not (mode == 'resident' and scenario == 'edge_write'):
    for k in edit:
        sp.p[k, 4] += np.float32(0.03125); p2['ux'][k] += np.float32(0.03125)
    if mode == 'resident':
        L.vpic_b200_invalidate.argtypes = [C.c_void_p]; L.vpic_b200_invalidate.restype = None
        L.vpic_b200_invalidate(sp.p.ctypes.data)
L.clear_accumulator_array(C.byref(H.aa))
L.advance_p(C.byref(sp.c), C.byref(H.aa), C.byref(H.ia))
if mode == 'resident': L.vpic_b200_sync_to_host(None)
pm2 = np.zeros(n, dtype=abi.mover_dtype); acc2 = np.zeros(((g.nv + 1)//2*2, 12), np.float32)
f32 = np.float32; dt = f32(g.dt)
a = R.OraclePushArgs(p2.ctypes.data, n, pm2.ctypes.data, n, interp.ctypes.data, 20, acc2.ctypes.data, 12,
                     g.neighbor.ctypes.data, g.rangel, g.rangeh, f32(f32(f32(-1)*dt)/f32(2)), dt, dt, dt, f32(-1))
orc.vpo_advance_p(C.byref(a), None)
got = sp.p[:n].reshape(-1).view(abi.particle_dtype)
bad = np.nonzero(np.any(got.view(np.uint32).reshape(n, 8) != p2.view(np.uint32).reshape(n, 8), axis=1))[0]
res = dict(seen=seen, particles=int(bad.size), first_bad=[int(x) for x in bad[:5]],
           partition=bool(np.array_equal(sp.partition[:g.nv], part[:g.nv])),
           accum=float(np.abs(H.accum - acc2).max() / np.abs(acc2).max()))
L.vpic_b200_set_mode(0)
import json; print('RESULT ' + json.dumps(res))
"""


_KEYS_SCRIPT = """
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import bench, refvpic as R
from vpic_b200 import lib, grid as G, abi
mode, scenario = sys.argv[1], sys.argv[2]
L = lib.load(); orc = R.load_oracle()
nx, ny, nz, n = 9, 8, 7, 70001
g = G.partition_periodic_box(0,0,0,nx,ny,nz,nx,ny,nz,1,1,1, dt=G.courant_dt(1,1,1,nx,ny,nz,frac=0.98))
H = bench.HostWorld(L, g, pinned='register')
L.vpic_b200_set_lazy_min.argtypes = [C.c_size_t]; L.vpic_b200_set_lazy_min.restype = None
L.vpic_b200_set_lazy_min(4096)
L.vpic_b200_sync_to_host.argtypes = [C.c_void_p]; L.vpic_b200_sync_to_host.restype = None
L.vpic_b200_set_mode(dict(resident=1, auto=2)[mode])
rng = np.random.default_rng(35)
fld = R.random_fields(rng, g.nv) * 0.05
H.fields[:] = fld
sp = H.new_species('e', -1.0, 1.0, n + 64, n, 3)
parts = R.random_particles(rng, n, nx, ny, nz, uth=0.3, w=0.5)
sp.p[:n] = parts.view(np.float32).reshape(-1, 8); sp.c.np = n
H.load_interpolator()
interp = np.zeros((g.nv, 20), np.float32)
orc.vpo_load_interpolator(interp.ctypes.data, 20, fld.ctypes.data, nx, ny, nz)
f32 = np.float32; dt = f32(g.dt)
def oracle_push(p2):
    pm2 = np.zeros(n, dtype=abi.mover_dtype); acc2 = np.zeros(((g.nv + 1)//2*2, 12), np.float32)
    a = R.OraclePushArgs(p2.ctypes.data, n, pm2.ctypes.data, n, interp.ctypes.data, 20, acc2.ctypes.data, 12,
                         g.neighbor.ctypes.data, g.rangel, g.rangeh, f32(f32(f32(-1)*dt)/f32(2)), dt, dt, dt, f32(-1))
    assert orc.vpo_advance_p(C.byref(a), None) == 0
def oracle_sort(p2):
    aux = np.zeros_like(p2); part = np.zeros(g.nv + 1, np.int32)
    orc.vpo_sort_p(p2.ctypes.data, n, aux.ctypes.data, part.ctypes.data, nx, ny, nz)
    return part
p2 = parts.copy()
# step 2 of a species sorted every 3 steps: this push knows that a sort opens the next step
H.G.step = 2
L.clear_accumulator_array(C.byref(H.aa)); L.advance_p(C.byref(sp.c), C.byref(H.aa), C.byref(H.ia))
oracle_push(p2)
other = int(p2['i'][n // 3])
if scenario == 'edge_edit' and mode == 'auto':   # the host moves a particle at the unprotected head of the array to another voxel
    view = sp.p.view(np.int32)
    view[2, 3] = other; p2['i'][2] = other
elif scenario == 'interior_edit' and mode == 'auto':      # ... and one in the middle (faults the chunk back)
    view = sp.p.view(np.int32)
    view[n // 2, 3] = other; p2['i'][n // 2] = other
H.G.step = 3
L.sort_p(C.byref(sp.c))
part = oracle_sort(p2)
L.clear_accumulator_array(C.byref(H.aa)); L.advance_p(C.byref(sp.c), C.byref(H.aa), C.byref(H.ia))
oracle_push(p2)
if mode == 'resident': L.vpic_b200_sync_to_host(None)
got = sp.p[:n].reshape(-1).view(abi.particle_dtype)
bad = np.nonzero(np.any(got.view(np.uint32).reshape(n, 8) != p2.view(np.uint32).reshape(n, 8), axis=1))[0]
res = dict(particles=int(bad.size), first_bad=[int(x) for x in bad[:5]], partition=bool(np.array_equal(sp.partition[:g.nv], part[:g.nv])))
L.vpic_b200_set_mode(0)
import json; print('RESULT ' + json.dumps(res))
"""


@pytest.mark.parametrize("mode", ["auto", "resident"])
@pytest.mark.parametrize("scenario", ["plain", "edge_edit", "interior_edit"])
def test_dropin_sort_p_on_the_keys_of_the_last_push(eng, oracle, mode, scenario):
    """The drop-in advance_p of the step before a sort (g->step + 1 divisible by sp->sort_interval) leaves the voxel
    keys behind and sort_p sorts those instead of reading the particles — unless the host touched the array in between:
    an edit at the unprotected head is picked up by re-reading those keys, an edit that faulted a chunk back voids the
    short cut.  Result after push, sort, push bit-identical to the oracle with the same edits."""
    import subprocess, sys, os, json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VPIC_B200_LAZY_CHUNK="16384", VPIC_B200_TRACE="1")
    r = subprocess.run([sys.executable, "-c", _KEYS_SCRIPT.format(root=root), mode, scenario],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][0][7:])
    assert res["particles"] == 0 and res["partition"], res
    trace = [ln for ln in r.stderr.splitlines() if "sort_p_on_keys_of_the_last_push" in ln][0]
    with_keys = int(trace.split("sort_p_on_keys_of_the_last_push=")[1].split()[0])
    assert with_keys == (0 if (scenario == "interior_edit" and mode == "auto") else 1), trace


@pytest.mark.parametrize("mode", ["auto", "resident"])
@pytest.mark.parametrize("scenario", ["plain", "edge_read", "edge_write", "interior_read"])
def test_dropin_deferred_sort_is_invisible_to_the_host(eng, oracle, mode, scenario):
    """sort_p of the drop-in layer only computes the order; the next advance_p moves the particles.  A host that reads
    or edits particles in between (collision operators do, advance.cc:44-47) must still see and get the sorted array:
    reads at the unprotected ends of the array, reads that fault a chunk back, edits at the ends that the fused push
    has to pick up in sorted positions.  Result after the push bit-identical to the oracle's sort -> edit -> push."""
    import subprocess, sys, os, json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VPIC_B200_LAZY_CHUNK="16384", VPIC_B200_TRACE="1")
    r = subprocess.run([sys.executable, "-c", _DEFER_SCRIPT.format(root=root), mode, scenario],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][0][7:])
    assert res["particles"] == 0 and res["partition"] and res["accum"] < 2e-5, res
    assert all(res["seen"].values()), res
    trace = [ln for ln in r.stderr.splitlines() if "sort_p_fused_into_advance_p" in ln][0]
    fused = int(trace.split("sort_p_fused_into_advance_p=")[1].split()[0])
    settled = int(trace.split("sort_p_applied_separately=")[1].split()[0])
    if scenario == "interior_read" or (mode == "resident" and scenario == "edge_read"):
        assert (fused, settled) == (0, 1), trace           # the host looked: the order was applied on its own
    else:
        assert (fused, settled) == (1, 0), trace


_AUTO_SCRIPT = """
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import bench, refvpic as R
from vpic_b200 import lib, grid as G, abi
mode, out = sys.argv[1], sys.argv[2]
L = lib.load()
nx, ny, nz, n = 9, 8, 7, 70001
g = G.partition_periodic_box(0,0,0,nx,ny,nz,nx,ny,nz,1,1,1, dt=G.courant_dt(1,1,1,nx,ny,nz,frac=0.98))
H = bench.HostWorld(L, g, pinned='register')
L.vpic_b200_set_lazy_min.argtypes = [C.c_size_t]; L.vpic_b200_set_lazy_min.restype = None
L.vpic_b200_set_lazy_min(4096)
L.vpic_b200_set_mode(dict(coherent=0, auto=2)[mode])
rng = np.random.default_rng(21)
H.fields[:] = R.random_fields(rng, g.nv) * 0.05
sp = H.new_species('e', -1.0, 1.0, n + 64, n, 3)
parts = R.random_particles(rng, n, nx, ny, nz, uth=0.3, w=0.5)
sp.p[:n] = parts.view(np.float32).reshape(-1, 8); sp.c.np = n
H.load_interpolator()
energies = []
libc = C.CDLL(None)
libc.fopen.restype = C.c_void_p; libc.fopen.argtypes = [C.c_char_p, C.c_char_p]; libc.fclose.argtypes = [C.c_void_p]
L.fwrite.restype = C.c_size_t; L.fwrite.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
aa, ia, fa = C.byref(H.aa), C.byref(H.ia), C.byref(H.fa)
for step in range(7):
    # particle side of the step only: E and B stay what the host set, so both runs push bit-identically
    if step % 3 == 0:
        L.sort_p(C.byref(sp.c))
    L.clear_accumulator_array(aa)
    L.advance_p(C.byref(sp.c), aa, ia)
    L.reduce_accumulator_array(aa)
    L.vpic_b200_clear_jf(fa)
    L.unload_accumulator_array(fa, aa)
    energies.append(L.energy_p(C.byref(sp.c), C.byref(H.ia)))
    # what decks do between steps: poke single particles, read a few, append one (inject_particle), touch a field
    k = (step * 9973) % n
    sp.p[k, 4] += np.float32(0.01)                                  # host write into a device-owned chunk
    energies.append(float(sp.p[(k * 7) % n, 5]))                    # host read
    if step == 2:
        sp.p[sp.c.np] = sp.p[0]; sp.p[sp.c.np, 0] = 0.25; sp.c.np += 1    # grow the live extent by one particle
    if step == 4:
        H.fields[g.nv // 2, 0] += np.float32(0.125)                 # set_region_field-like
        H.load_interpolator()
    if step == 5:                                                   # dump: fwrite of a device-owned array (no fault: syscall)
        f = libc.fopen(out.encode() + b'.dump', b'wb')
        wrote = L.fwrite(sp.p.ctypes.data, 32, sp.c.np, f); libc.fclose(f)
        assert wrote == sp.c.np, wrote
st = (C.c_uint64 * 4)(); L.vpic_b200_lazy_stats(st)
tb = H.transfer_bytes()
np.savez(out, p=sp.p[:sp.c.np].copy(), f=H.fields.copy(), i=H.interp.copy(), a=H.accum.copy(), e=np.array(energies),
         lazy=np.array([int(x) for x in st]), tb=np.array(tb, dtype=np.float64),
         dump=np.fromfile(out + '.dump', dtype=np.float32))
L.vpic_b200_set_mode(0)
"""


def test_dropin_auto_mode_matches_coherent_under_host_interference(eng, tmp_path):
    """VPB_MODE_AUTO must be indistinguishable from VPB_MODE_COHERENT for a host that reads, writes, grows and dumps
    its arrays between calls — while moving far fewer bytes.  Same script twice, results compared bit for bit
    (currents are summed with atomics in no fixed order, so they are compared with a tolerance)."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ("coherent", "auto"):
        out = str(tmp_path / f"{mode}.npz")
        env = dict(os.environ, VPIC_B200_LAZY_CHUNK="16384")
        r = subprocess.run([sys.executable, "-c", _AUTO_SCRIPT.format(root=root), mode, out],
                           capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        res[mode] = np.load(out)
    c, a = res["coherent"], res["auto"]
    assert c["p"].shape == a["p"].shape
    # E and B are never advanced in the script, so the push is bit-reproducible; only the currents (atomic order) are not
    assert np.array_equal(c["p"].view(np.uint32), a["p"].view(np.uint32))
    assert np.array_equal(c["i"].view(np.uint32), a["i"].view(np.uint32))
    assert np.array_equal(c["f"][:, :12].view(np.uint32), a["f"][:, :12].view(np.uint32))
    np.testing.assert_allclose(a["f"], c["f"], rtol=0, atol=2e-5 * max(1.0, float(np.abs(c["f"]).max())))
    np.testing.assert_allclose(a["a"], c["a"], rtol=0, atol=2e-5 * max(1.0, float(np.abs(c["a"]).max())))
    np.testing.assert_allclose(a["e"], c["e"], rtol=1e-6, atol=0)
    assert a["dump"].size == c["dump"].size > 0
    assert np.array_equal(a["dump"].view(np.uint32), c["dump"].view(np.uint32))   # the dump saw current data, not stale pages
    faults, fault_bytes, remaps, regions = a["lazy"]
    assert faults >= 10 and regions >= 5 and remaps == 0
    assert c["lazy"][0] == 0
    assert a["tb"].sum() < 0.5 * c["tb"].sum(), (a["tb"], c["tb"])             # and it moved much less data


def test_accumulate_rho_p(eng, oracle):
    rng = np.random.default_rng(12)
    nx, ny, nz = 9, 7, 5
    g = make_grid(nx, ny, nz)
    dg = eng.DeviceGrid(g)
    fields = R.random_fields(rng, g.nv)
    fa = eng.FieldArray(dg)
    fa.f.copy_(torch.from_numpy(fields))
    parts = R.random_particles(rng, 50001, nx, ny, nz, w=0.21)
    sp = eng.Species("ion", 2.0, 25.0, len(parts), 16, 20, 0, dg)
    sp.set_particles(parts)
    eng.accumulate_rho_p(fa, sp)
    f_ref = fields.copy()
    oracle.vpo_accumulate_rho_p(f_ref.ctypes.data, parts.ctypes.data, len(parts), 2.0, g.r8V, nx, ny, nz)
    got = fa.f.cpu().numpy()
    assert np.array_equal(bits(np.delete(got, 15, axis=1)), bits(np.delete(f_ref, 15, axis=1)))     # only rhof changes
    assert np.abs(got[:, 15] - f_ref[:, 15]).max() / np.abs(f_ref[:, 15]).max() < 2e-5


def test_boundary_p_absorbing_walls_match_reference(eng, oracle):
    """Absorbing walls on one rank: movers -> device boundary_p (back-fill + rhob) vs the reference's own boundary_p
    (when the prebuilt reference is present) or its restated semantics (sequential p[i] = p[--np])."""
    rng = np.random.default_rng(6)
    nx, ny, nz, n = 6, 5, 4, 12000
    pbc = {0: -2, 3: -2, 2: -2}
    g = make_grid(nx, ny, nz, pbc=pbc)
    fields = R.random_fields(rng, g.nv)
    interp = np.zeros((g.nv, 20), dtype=np.float32)
    oracle.vpo_load_interpolator(interp.ctypes.data, 20, fields.ctypes.data, nx, ny, nz)
    parts = R.random_particles(rng, n, nx, ny, nz, uth=0.5, w=0.7)
    # expected result from the oracle push followed by the reference's sequential removal
    p_ref, pm_ref, acc_ref, _ = oracle_push(oracle, g, parts, interp, -1.0, 1.0, n)
    assert len(pm_ref) > 50
    f_ref = fields.copy()
    np_ref = n
    for m in pm_ref[::-1]:                                  # boundary_p.cc:257-371: reverse walk, back-fill
        i = int(m["i"])
        one = p_ref[i:i + 1].copy()
        one["i"] >>= 3
        oracle.vpo_accumulate_rhob(f_ref.ctypes.data, one.ctypes.data, -1.0, g.r8V, nx, ny, nz)
        np_ref -= 1
        p_ref[i] = p_ref[np_ref]
    dg = eng.DeviceGrid(g)
    fa, ia, aa = eng.FieldArray(dg), eng.InterpolatorArray(dg), eng.AccumulatorArray(dg)
    fa.f.copy_(torch.from_numpy(fields)); ia.i.copy_(torch.from_numpy(interp))
    sp = eng.Species("e", -1.0, 1.0, n, n, 20, 0, dg)
    sp.set_particles(parts)
    eng.advance_p(sp, aa, ia)
    assert sp.nm == len(pm_ref)
    inj, offs = eng.boundary_pack(sp, [-1] * 6, fa)
    o = offs.cpu().numpy()
    assert o[6] == 0 and o[7] == len(pm_ref) and o[8] == len(pm_ref)          # every mover was absorbed
    assert sp.np == np_ref and sp.nm == 0
    assert np.array_equal(bits(sp.particles_host()), bits(p_ref[:np_ref])), "back-fill must equal the sequential one"
    got = fa.f.cpu().numpy()
    assert np.abs(got[:, 11] - f_ref[:, 11]).max() <= 2e-5 * np.abs(f_ref[:, 11]).max()
    assert np.array_equal(bits(np.delete(got, 11, axis=1)), bits(np.delete(f_ref, 11, axis=1)))
    if R.have_ref("scalar"):                               # and against the unmodified reference's boundary_p itself
        lib = R.load_ref("scalar", tpp=1)
        W = R.RefWorld(lib, nx, ny, nz, pbc=pbc)
        W.g.contents.dt = g.dt
        W.fields[:] = fields
        lib.load_interpolator_array(W.ia, W.fa)
        rs = W.new_species("e_bp", -1.0, 1.0, n, n)
        rs.set_particles(parts)
        lib.clear_accumulator_array(W.aa); lib.advance_p(rs.sp, W.aa, W.ia); lib.reduce_accumulator_array(W.aa)
        lib.boundary_p.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.boundary_p(None, rs.sp, W.fa, W.aa)
        assert rs.c.np == sp.np
        assert np.array_equal(bits(rs.p[:rs.c.np]), bits(sp.particles_host()))
        assert np.abs(got[:, 11] - W.fields[:, 11]).max() <= 2e-5 * np.abs(W.fields[:, 11]).max()


_BOUNDARY_SCRIPT = """
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import bench, refvpic as R
from vpic_b200 import lib, grid as G, abi
mode = sys.argv[1]
L = lib.load()
nx, ny, nz, n = 6, 5, 4, 12000
pbc = {{0: -2, 3: -2, 2: -2}}
g = G.partition_periodic_box(0,0,0,nx,ny,nz,nx,ny,nz,1,1,1, dt=G.courant_dt(1,1,1,nx,ny,nz,frac=0.98))
for f, code in pbc.items():
    g.set_pbc(f, code)
H = bench.HostWorld(L, g, pinned='register')
L.vpic_b200_set_lazy_min.argtypes = [C.c_size_t]; L.vpic_b200_set_lazy_min.restype = None
L.vpic_b200_set_lazy_min(4096)
L.vpic_b200_set_mode(dict(coherent=0, auto=2)[mode])
rng = np.random.default_rng(6)
fields = R.random_fields(rng, g.nv)
H.fields[:] = fields
sp = H.new_species('e', -1.0, 1.0, n, n, 20)
parts = R.random_particles(rng, n, nx, ny, nz, uth=0.5, w=0.7)
sp.p[:n] = parts.view(np.float32).reshape(-1, 8); sp.c.np = n
H.load_interpolator()
L.clear_accumulator_array(C.byref(H.aa))
L.advance_p(C.byref(sp.c), C.byref(H.aa), C.byref(H.ia))
L.reduce_accumulator_array(C.byref(H.aa))
nm = int(sp.c.nm)
L.boundary_p.argtypes = [C.c_void_p] * 4; L.boundary_p.restype = None
L.boundary_p(None, C.byref(sp.c), C.byref(H.fa), C.byref(H.aa))
# the unmodified reference on the same inputs
ref = R.load_ref('scalar', tpp=1)
W = R.RefWorld(ref, nx, ny, nz, pbc=pbc)
W.g.contents.dt = g.dt
W.fields[:] = fields
ref.load_interpolator_array(W.ia, W.fa)
rs = W.new_species('e_bp', -1.0, 1.0, n, n)
rs.set_particles(parts)
ref.clear_accumulator_array(W.aa); ref.advance_p(rs.sp, W.aa, W.ia); ref.reduce_accumulator_array(W.aa)
ref.boundary_p.argtypes = [C.c_void_p] * 4
ref.boundary_p(None, rs.sp, W.fa, W.aa)
got = sp.p[:sp.c.np].copy(); want = rs.p[:rs.c.np]
rhob, rhob_ref = H.fields[:, 11].copy(), W.fields[:, 11].copy()
others = bool(np.array_equal(np.delete(H.fields, [11, 12, 13, 14], axis=1).view(np.uint32),
                             np.delete(W.fields, [11, 12, 13, 14], axis=1).view(np.uint32)))
import json
print('RESULT ' + json.dumps(dict(nm=nm, np=int(sp.c.np), np_ref=int(rs.c.np), nm_after=int(sp.c.nm),
      particles=bool(got.nbytes == want.nbytes and np.array_equal(got.view(np.uint8).ravel(), want.view(np.uint8).ravel())),
      rhob=float(np.abs(rhob - rhob_ref).max() / np.abs(rhob_ref).max()), others=others)))
L.vpic_b200_set_mode(0)
"""


@pytest.mark.parametrize("mode", ["coherent", "auto"])
def test_dropin_boundary_p_absorbing_walls(eng, mode):
    """boundary_p(particle_bc_t*, species_t*, field_array_t*, accumulator_array_t*) on HOST structs, one rank, three
    absorbing walls: the particle array after the device back-fill must be bit-identical to the one the unmodified
    reference's boundary_p leaves, and the absorbed charge in rhob must agree to summation-order tolerance."""
    import subprocess, sys, os, json
    if not R.have_ref("scalar"):
        pytest.skip("oracle/_ref not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VPIC_B200_LAZY_CHUNK="8192")
    r = subprocess.run([sys.executable, "-c", _BOUNDARY_SCRIPT.format(root=root), mode],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][0][7:])
    assert res["nm"] > 50 and res["nm_after"] == 0 and res["np"] == res["np_ref"] == 12000 - res["nm"], res
    assert res["particles"], "particle array differs from the reference's after boundary_p"
    assert res["others"], "boundary_p touched field slots other than rhob"
    assert res["rhob"] < 2e-5, res["rhob"]


def test_dropin_move_p_runs_on_the_device_when_the_arrays_live_there(eng, oracle):
    """move_p(particle_t*, particle_mover_t*, accumulator_t*, const grid_t*, qsp) — species_advance.h:152-157 — is what
    inject_particle and the emitters call on single particles.  In resident mode, after an advance_p has left the
    arrays on the device, the exported symbol must move the particle THERE (no page of the particle array comes back)
    and produce the reference's result: same particle bytes, same mover, same return value, same deposits."""
    import subprocess, sys, os, json, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent(f"""
        import sys, json, ctypes as C, numpy as np
        sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})
        import bench, refvpic as R
        from vpic_b200 import lib, grid as G, abi
        L = lib.load(); orc = R.load_oracle()
        nx, ny, nz, n = 7, 6, 5, 5003
        g = G.partition_periodic_box(0,0,0,nx,ny,nz,nx,ny,nz,1,1,1, dt=G.courant_dt(1,1,1,nx,ny,nz,frac=0.98))
        g.set_pbc(2, -2); g.set_pbc(5, -1)                      # one absorbing and one reflecting z wall
        H = bench.HostWorld(L, g, pinned=True)
        L.vpic_b200_set_mode(1)                                  # resident
        rng = np.random.default_rng(12)
        sp = H.new_species('e', -1.0, 1.0, n, n, 20)
        parts = R.random_particles(rng, n, nx, ny, nz, uth=0.0, w=0.5)
        sp.p[:n] = parts.view(np.float32).reshape(-1, 8); sp.c.np = n
        H.load_interpolator()
        L.clear_accumulator_array(C.byref(H.aa))
        L.advance_p(C.byref(sp.c), C.byref(H.aa), C.byref(H.ia))        # zero momenta, zero fields: nothing moves
        L.vpic_b200_sync_to_host.argtypes = [C.c_void_p]; L.vpic_b200_sync_to_host.restype = None
        L.vpic_b200_transfer_bytes.argtypes = [C.c_void_p]
        L.vpic_b200_sync_to_host(None)
        p_ref = sp.p[:n].reshape(-1).view(abi.particle_dtype).copy()
        acc_ref = H.accum.copy()
        # the arrays are on the device again after the next hot-path call
        L.clear_accumulator_array(C.byref(H.aa)); acc_ref[:] = 0
        L.advance_p(C.byref(sp.c), C.byref(H.aa), C.byref(H.ia))
        L.move_p.restype = C.c_int
        L.move_p.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        before = (C.c_uint64 * 2)(); L.vpic_b200_transfer_bytes(before)
        moves = []
        for k in range(40):
            mv = np.zeros(1, dtype=abi.mover_dtype)
            mv['i'] = int(rng.integers(0, n))
            mv['dispx'], mv['dispy'], mv['dispz'] = (rng.uniform(-0.9, 0.9, 3)).astype(np.float32)
            mv2 = mv.copy()
            ret = L.move_p(sp.p.ctypes.data, mv.ctypes.data, H.accum.ctypes.data, C.byref(H.G), C.c_float(-1.0))
            ret2 = orc.vpo_move_p(p_ref.ctypes.data, mv2.ctypes.data, acc_ref.ctypes.data, 12, g.neighbor.ctypes.data,
                                  g.rangel, g.rangeh, C.c_float(-1.0))
            moves.append(bool(ret == ret2 and np.array_equal(mv.view(np.uint8), mv2.view(np.uint8))))
        after = (C.c_uint64 * 2)(); L.vpic_b200_transfer_bytes(after)
        L.vpic_b200_sync_to_host(None)
        got = sp.p[:n].reshape(-1).view(abi.particle_dtype)
        ok = dict(moves=all(moves), particles=bool(np.array_equal(got.view(np.uint8), p_ref.view(np.uint8))),
                  accum=float(np.abs(H.accum - acc_ref).max() / max(np.abs(acc_ref).max(), 1e-30)),
                  d2h_during_moves=int(after[1] - before[1]), changed=int((got['i'] != parts['i']).sum()))
        L.vpic_b200_set_mode(0)
        print('RESULT ' + json.dumps(ok))
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][0][7:])
    assert res["moves"] and res["particles"] and res["accum"] < 1e-5, res
    assert res["changed"] > 5                                    # particles really crossed cells
    assert res["d2h_during_moves"] < 40 * 64, res                # only movers and return values crossed PCIe

"""GPU parity of the brick/tile advance_p (csrc/advance_p_brick.cu) against the CPU oracle.

The brick kernel is what vpb_advance_p runs once a species has been sorted (it needs partition[]).  Particle bytes and
movers must be bit-exact; accumulators agree to summation-order tolerance.  Cases: every kernel configuration
(packed / scalar arithmetic, movers into the tile or to global memory, the tile geometries), several steps of drift
after one sort (particles leave their bricks and tiles), walls, thin grids, a stale partition (particles appended or
removed since the sort), and one run at the benchmark's scale (> 100 M particles, multi-span warps and the grid cap).
"""
import os
import numpy as np
import pytest

import refvpic as R
from test_gpu_parity import make_grid, oracle_push, accum_close, bits

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vpic_b200 import engine, lib
    lib.load()
    return engine


def setup(eng, oracle, dims, n, uth, pbc, seed):
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz, pbc=pbc)
    fields = R.random_fields(rng, g.nv)
    interp = np.zeros((g.nv, 20), dtype=np.float32)
    oracle.vpo_load_interpolator(interp.ctypes.data, 20, fields.ctypes.data, nx, ny, nz)
    parts = R.random_particles(rng, n, nx, ny, nz, uth=uth, w=0.7)
    dg = eng.DeviceGrid(g)
    ia, aa = eng.InterpolatorArray(dg), eng.AccumulatorArray(dg)
    ia.i.copy_(torch.from_numpy(interp))
    return g, dg, ia, aa, interp, parts


def check_step(eng, oracle, g, sp, aa, ia, interp, max_nm, q=-1.0, m=1.0):
    p_in = sp.particles_host().copy()
    p_ref, pm_ref, acc_ref, _ = oracle_push(oracle, g, p_in, interp, q, m, max_nm)
    eng.clear_accumulator_array(aa)
    eng.advance_p(sp, aa, ia, variant=5)
    assert sp.nm == len(pm_ref) and sp.n_ignored == 0
    got = sp.particles_host()
    assert np.array_equal(bits(got), bits(p_ref)), "particle state must be bit-exact"
    assert np.array_equal(bits(sp.movers_host()), bits(pm_ref)), "movers must be bit-exact and ascending"
    accum_close(aa.a.cpu().numpy(), acc_ref)
    return got


BRICK_CASES = [
    ((12, 9, 7), 40000, 0.5, None, 4),            # periodic, fast particles: leave bricks and tiles within a few steps
    ((16, 1, 16), 30011, 0.3, None, 3),           # 2-D deck shape (one cell in y), ragged np
    ((8, 6, 5), 9000, 0.6, {0: -1, 3: -1}, 3),    # reflecting x walls
    ((9, 4, 4), 9000, 0.6, {2: -2, 5: -2}, 1),    # absorbing z walls -> movers are emitted
    ((6, 6, 6), 1, 0.1, None, 1),                 # single particle
    ((40, 8, 8), 200000, 0.1, None, 2),           # slow particles, many rows per voxel: the summed (sorted) path
]


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("dims,n,uth,pbc,steps", BRICK_CASES)
def test_advance_p_brick(eng, oracle, monkeypatch, dims, n, uth, pbc, steps, cfg):
    monkeypatch.setenv("VPB_BRICK_CFG", str(cfg | 0x100))
    g, dg, ia, aa, interp, parts = setup(eng, oracle, dims, n, uth, pbc, 41)
    max_nm = max(16, n)
    sp = eng.Species("electron", -1.0, 1.0, max(n, 1), max_nm, 20, 0, dg)
    sp.set_particles(parts)
    eng.sort_p(sp)
    before = eng._lib.load().vpb_launch_count()
    for _ in range(steps):
        got = check_step(eng, oracle, g, sp, aa, ia, interp, max_nm)
        if sp.nm:
            break
    assert eng._lib.load().vpb_launch_count() > before


@pytest.mark.parametrize("cfg", [0, 5])
def test_advance_p_brick_stale_partition(eng, oracle, monkeypatch, cfg):
    """The array changed after the sort: particles appended (boundary_p injection) and removed (back-fill).  The bricks'
    segments plus the tail must still cover every particle exactly once."""
    monkeypatch.setenv("VPB_BRICK_CFG", str(cfg | 0x100))
    dims, n = (10, 6, 6), 30000
    g, dg, ia, aa, interp, parts = setup(eng, oracle, dims, n, 0.3, None, 5)
    extra = R.random_particles(np.random.default_rng(77), 9000, *dims, uth=0.3, w=0.7)
    sp = eng.Species("electron", -1.0, 1.0, n + len(extra), n + len(extra), 20, 0, dg)
    sp.set_particles(parts)
    eng.sort_p(sp)
    check_step(eng, oracle, g, sp, aa, ia, interp, sp.max_nm)
    # append: np grows beyond partition[nv]
    sp.p[sp.np:sp.np + len(extra)].copy_(torch.from_numpy(extra.view(np.float32).reshape(-1, 8)))
    sp.np += len(extra)
    check_step(eng, oracle, g, sp, aa, ia, interp, sp.max_nm)
    # remove: np falls below partition[nv]; the last segments are clipped
    sp.np = n - 7001
    check_step(eng, oracle, g, sp, aa, ia, interp, sp.max_nm)
    # and an ion species (positive charge, heavier) through the same path
    sp2 = eng.Species("ion", 1.0, 25.0, n, n, 20, 0, dg)
    sp2.set_particles(parts)
    eng.sort_p(sp2)
    check_step(eng, oracle, g, sp2, aa, ia, interp, n, q=1.0, m=25.0)


def test_advance_p_brick_accum_stride16(eng, oracle, monkeypatch):
    """accumulator_t padded to 16 floats (V8/V16 host builds)."""
    monkeypatch.setenv("VPB_BRICK_CFG", str(0x100))
    dims, n = (10, 6, 6), 30000
    rng = np.random.default_rng(3)
    nx, ny, nz = dims
    g = make_grid(nx, ny, nz)
    fields = R.random_fields(rng, g.nv)
    interp = np.zeros((g.nv, 24), dtype=np.float32)
    oracle.vpo_load_interpolator(interp.ctypes.data, 24, fields.ctypes.data, nx, ny, nz)
    parts = R.random_particles(rng, n, nx, ny, nz, uth=0.4, w=0.7)
    dg = eng.DeviceGrid(g)
    ia, aa = eng.InterpolatorArray(dg, simd_width=8), eng.AccumulatorArray(dg, simd_width=8)
    assert ia.stride == 24 and aa.stride_floats == 16
    ia.i.copy_(torch.from_numpy(interp))
    sp = eng.Species("electron", -1.0, 1.0, n, n, 20, 0, dg)
    sp.set_particles(parts)
    eng.sort_p(sp)
    for _ in range(2):
        p_in = sp.particles_host().copy()
        p_ref, pm_ref, acc_ref, _ = oracle_push(oracle, g, p_in, interp, -1.0, 1.0, n, isf=24, asf=16)
        eng.clear_accumulator_array(aa)
        eng.advance_p(sp, aa, ia, variant=5)
        assert np.array_equal(bits(sp.particles_host()), bits(p_ref))
        accum_close(aa.a.cpu().numpy(), acc_ref)


@pytest.mark.parametrize("variant", [5, 2])
def test_advance_p_at_scale(eng, oracle, variant):
    """> 100 M particles, the benchmark's regime: every warp takes several work items, the grid is capped at the
    machine size.  Brick kernel (variant 5) and the linear kernel (variant 2, multi-span warps) against the oracle."""
    dims, ppc = (128, 128, 100), 64
    nx, ny, nz = dims
    n = nx * ny * nz * ppc                                       # 104 857 600
    rng = np.random.default_rng(99)
    g = make_grid(nx, ny, nz)
    fields = R.random_fields(rng, g.nv)
    interp = np.zeros((g.nv, 20), dtype=np.float32)
    oracle.vpo_load_interpolator(interp.ctypes.data, 20, fields.ctypes.data, nx, ny, nz)
    dg = eng.DeviceGrid(g)
    ia, aa = eng.InterpolatorArray(dg), eng.AccumulatorArray(dg)
    ia.i.copy_(torch.from_numpy(interp))
    sp = eng.Species("electron", -1.0, 1.0, n, n // 8, 20, 0, dg)
    # particles generated on the device (the host generator takes minutes at this size); unique weights as tags
    gen = torch.Generator(device="cuda").manual_seed(5)
    p = sp.p[:n]
    p[:, 0:3] = torch.rand((n, 3), generator=gen, device="cuda") * 2 - 1
    ix = torch.randint(1, nx + 1, (n,), generator=gen, device="cuda", dtype=torch.int32)
    iy = torch.randint(1, ny + 1, (n,), generator=gen, device="cuda", dtype=torch.int32)
    iz = torch.randint(1, nz + 1, (n,), generator=gen, device="cuda", dtype=torch.int32)
    sp.p.view(torch.int32)[:n, 3] = ix + (nx + 2) * (iy + (ny + 2) * iz)
    p[:, 4:7] = torch.randn((n, 3), generator=gen, device="cuda") * 0.18
    p[:, 7] = torch.rand((n,), generator=gen, device="cuda") + 0.5
    del ix, iy, iz
    sp.np = n
    eng.sort_p(sp)
    if variant != 5:
        sp.partition_np = -1                                     # linear kernel
    p_in = sp.particles_host().copy()
    assert np.all(np.diff(p_in["i"]) >= 0), "sort_p at scale must order the voxels"
    max_nm = sp.max_nm
    p_ref, pm_ref, acc_ref, _ = oracle_push(oracle, g, p_in, interp, -1.0, 1.0, max_nm)
    del p_in
    eng.clear_accumulator_array(aa)
    eng.advance_p(sp, aa, ia, variant=variant)
    assert sp.nm == len(pm_ref) == 0
    got = sp.particles_host()
    assert np.array_equal(bits(got), bits(p_ref)), "particle state must be bit-exact at scale"
    accum_close(aa.a.cpu().numpy(), acc_ref)

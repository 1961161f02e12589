"""Golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py -> hotpath_golden.npz):
 * CPU (`-m "not gpu"`): the oracle restatement reproduces them bit-for-bit — this pins the oracle even where neither
   /root/reference nor oracle/_ref is present;
 * GPU (`-m gpu`): the CUDA path, through the C-ABI, against the same reference outputs directly.
"""
import ctypes as C
import os
import numpy as np
import pytest

import refvpic as R
from vpic_b200 import abi, grid as G

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_golden.npz"))
CASES = ["box3d", "walls2d"]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def case(name):
    g = {k.split(".", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + ".")}
    nx, ny, nz = (int(v) for v in g["dims"])
    c = g["consts"]
    grid = G.partition_periodic_box(0, 0, 0, nx * float(c[4]), ny * float(c[5]), nz * float(c[6]), nx, ny, nz, 1, 1, 1, dt=float(c[0]))
    grid.neighbor = g["neighbor"].copy()
    grid.bc = [int(b) for b in g["bc"]]
    for k, v in zip(("dt", "cvac", "eps0", None, "dx", "dy", "dz", "dV", "rdx", "rdy", "rdz", "r8V"), c):
        if k:
            setattr(grid, k, float(v))
    return g, grid, (nx, ny, nz), float(c[3])


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_golden(oracle, name):
    g, grid, (nx, ny, nz), damp = case(name)
    f = g["fields0"].copy()
    interp = np.zeros_like(g["interp"])
    oracle.vpo_load_interpolator(interp.ctypes.data, 20, f.ctypes.data, nx, ny, nz)
    assert np.array_equal(bits(interp), bits(g["interp"]))
    p = g["p0"].copy(); aux = np.zeros_like(p); part = np.zeros(grid.nv + 1, np.int32)
    oracle.vpo_sort_p(p.ctypes.data, len(p), aux.ctypes.data, part.ctypes.data, nx, ny, nz)
    assert np.array_equal(bits(p), bits(g["p_sorted"])) and np.array_equal(part[:grid.nv], g["partition"])
    e_p = oracle.vpo_energy_p(p.ctypes.data, len(p), interp.ctypes.data, 20, -1.0, 1.0, grid.dt, grid.cvac)
    assert e_p == float(g["energy_p"][0])
    pm = np.zeros(len(p), dtype=abi.mover_dtype)
    acc = np.zeros_like(g["accum"])
    f32 = np.float32
    dt, cvac = f32(grid.dt), f32(grid.cvac)
    a = R.OraclePushArgs(p.ctypes.data, len(p), pm.ctypes.data, len(p), interp.ctypes.data, 20, acc.ctypes.data, 12,
                         grid.neighbor.ctypes.data, grid.rangel, grid.rangeh,
                         f32(f32(f32(-1) * dt) / f32(f32(f32(2) * f32(1)) * cvac)),
                         f32(f32(cvac * dt) * f32(grid.rdx)), f32(f32(cvac * dt) * f32(grid.rdy)),
                         f32(f32(cvac * dt) * f32(grid.rdz)), f32(-1))
    nm = oracle.vpo_advance_p(C.byref(a), None)
    assert np.array_equal(bits(p), bits(g["p1"]))
    assert nm == len(g["movers"]) and np.array_equal(bits(pm[:nm]), bits(g["movers"]))
    assert np.array_equal(bits(acc), bits(g["accum"]))
    fa = R.OracleFieldArgs()
    fa.f = f.ctypes.data; fa.nx, fa.ny, fa.nz = nx, ny, nz
    fa.dt, fa.cvac, fa.eps0, fa.damp = grid.dt, grid.cvac, grid.eps0, damp
    fa.dx, fa.dy, fa.dz, fa.dV = grid.dx, grid.dy, grid.dz, grid.dV
    fa.rdx, fa.rdy, fa.rdz = grid.rdx, grid.rdy, grid.rdz
    for i, (fi, fj, fk) in enumerate(G.FACES):
        fa.bc6[i] = grid.bc[G.boundary_index(fi, fj, fk)]
    oracle.vpo_clear_jf(C.byref(fa))
    oracle.vpo_unload_accumulator(f.ctypes.data, acc.ctypes.data, 12, nx, ny, nz, grid.rdx, grid.rdy, grid.rdz, grid.dt)
    oracle.vpo_synchronize_jf(C.byref(fa))
    assert np.array_equal(bits(f), bits(g["fields_jf"]))
    oracle.vpo_advance_b(C.byref(fa), 0.5); oracle.vpo_vacuum_advance_e(C.byref(fa), 1.0); oracle.vpo_advance_b(C.byref(fa), 0.5)
    assert np.array_equal(bits(f), bits(g["fields1"]))
    en = (C.c_double * 6)()
    oracle.vpo_vacuum_energy_f(C.byref(fa), en)
    assert np.array_equal(np.array(en[:]), g["energy_f"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_reference_golden(name):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vpic_b200 import engine as E
    g, grid, (nx, ny, nz), damp = case(name)
    dg = E.DeviceGrid(grid)
    fa, ia, aa = E.FieldArray(dg, damp=damp), E.InterpolatorArray(dg), E.AccumulatorArray(dg)
    fa.f.copy_(torch.from_numpy(g["fields0"].copy()))
    E.load_interpolator_array(ia, fa)
    assert np.array_equal(bits(ia.i.cpu().numpy()), bits(g["interp"]))
    n = len(g["p0"])
    sp = E.Species("electron", -1.0, 1.0, n, n, 20, 0, dg)
    sp.set_particles(g["p0"].copy())
    E.sort_p(sp)
    assert np.array_equal(bits(sp.particles_host()), bits(g["p_sorted"]))
    assert np.array_equal(sp.partition.cpu().numpy()[:grid.nv], g["partition"])
    assert abs(E.energy_p(sp, ia) - float(g["energy_p"][0])) <= 1e-12 * abs(float(g["energy_p"][0]))
    E.clear_accumulator_array(aa)
    E.advance_p(sp, aa, ia)
    E.reduce_accumulator_array(aa)
    assert np.array_equal(bits(sp.particles_host()), bits(g["p1"])), "particle state vs the reference must be bit-exact"
    assert sp.nm == len(g["movers"]) and np.array_equal(bits(sp.movers_host()), bits(g["movers"]))
    acc = aa.a.cpu().numpy()
    assert np.abs(acc - g["accum"]).max() <= 2e-5 * np.abs(g["accum"]).max()        # fp32 atomic order
    # continue from the reference's accumulators so the field comparison stays bit-exact
    aa.a.copy_(torch.from_numpy(g["accum"].copy()))
    fa.clear_jf(); E.unload_accumulator_array(fa, aa); fa.synchronize_jf()
    assert np.array_equal(bits(fa.f.cpu().numpy()), bits(g["fields_jf"]))
    fa.advance_b(0.5); fa.advance_e(1.0); fa.advance_b(0.5)
    assert np.array_equal(bits(fa.f.cpu().numpy()), bits(g["fields1"]))
    np.testing.assert_allclose(fa.energy_f(), g["energy_f"], rtol=1e-12)

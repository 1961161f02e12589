"""Generates tests/golden/hotpath_golden.npz by running the UNMODIFIED reference (oracle/_ref/libvpic_ref_scalar.so,
built from /root/reference by oracle/Makefile) on small seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The fixture lets the oracle and the CUDA path be checked against real reference outputs where neither
/root/reference nor oracle/_ref exists.  Cases: a periodic 3-D box, and a 2-D box (ny = 1) with reflecting x walls
and absorbing z walls (movers).  Sequence per case: load_interpolator_array, sort_p, clear_accumulator_array,
advance_p, reduce_accumulator_array, unload_accumulator_array, synchronize_jf, advance_b(1/2), advance_e, advance_b(1/2),
energy_p, energy_f.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refvpic as R   # noqa: E402

lib = R.load_ref("scalar", tpp=1)
out = {}
cases = [("box3d", (6, 5, 4), None, None, 0.45, 4000, 0.0),
         ("walls2d", (10, 1, 7), {0: -1, 3: -1, 2: -2, 5: -2}, {0: -1, 3: -1, 2: -1, 5: -1}, 0.5, 3008, 0.01)]
for name, (nx, ny, nz), pbc, fbc, uth, n, damp in cases:
    rng = np.random.default_rng(2024)
    W = R.RefWorld(lib, nx, ny, nz, pbc=pbc, fbc=fbc, damp=damp)
    g = W.g.contents
    W.fields[:] = R.random_fields(rng, W.nv)
    out[f"{name}.dims"] = np.array([nx, ny, nz], np.int32)
    out[f"{name}.consts"] = np.array([g.dt, g.cvac, g.eps0, damp, g.dx, g.dy, g.dz, g.dV, g.rdx, g.rdy, g.rdz, g.r8V], np.float32)
    out[f"{name}.bc"] = np.array(list(g.bc), np.int32)
    out[f"{name}.neighbor"] = W.neighbor.copy()
    out[f"{name}.fields0"] = W.fields.copy()
    lib.load_interpolator_array(W.ia, W.fa)
    out[f"{name}.interp"] = W.interp.copy()
    sp = W.new_species(f"gold_{name}", -1.0, 1.0, n, n)
    parts = R.random_particles(rng, n, nx, ny, nz, uth=uth, w=0.37)
    out[f"{name}.p0"] = parts.copy()
    sp.set_particles(parts)
    lib.sort_p(sp.sp)
    out[f"{name}.p_sorted"] = sp.p[:n].copy()
    out[f"{name}.partition"] = sp.partition[:W.nv].copy()
    lib.clear_accumulator_array(W.aa)
    lib.advance_p(sp.sp, W.aa, W.ia)
    lib.reduce_accumulator_array(W.aa)
    out[f"{name}.p1"] = sp.p[:n].copy()
    out[f"{name}.movers"] = sp.pm[:sp.c.nm].copy()
    out[f"{name}.accum"] = W.accum[0].copy()
    W.clear_jf()
    lib.unload_accumulator_array(W.fa, W.aa)
    W.synchronize_jf()
    out[f"{name}.fields_jf"] = W.fields.copy()
    W.advance_b(0.5); W.advance_e(1.0); W.advance_b(0.5)
    out[f"{name}.fields1"] = W.fields.copy()
    out[f"{name}.energy_f"] = W.energy_f()
    # energy_p needs in-domain voxel indices: measure it on the sorted (pre-push) particles
    sp.set_particles(out[f"{name}.p_sorted"])
    out[f"{name}.energy_p"] = np.array([lib.energy_p(sp.sp, W.ia)])
np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **out)
print("wrote hotpath_golden.npz", {k: v.shape for k, v in out.items() if k.startswith("box3d")})

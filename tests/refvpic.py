"""TEST INFRASTRUCTURE: drive the UNMODIFIED reference (oracle/_ref/libvpic_ref_*.so)
and the C restatement (oracle/liboracle.so) from Python through ctypes.

Nothing in vpic_b200/ imports this module.  The reference library is built by
oracle/Makefile from /root/reference in the build container and travels to the
GPU box as a prebuilt .so; tests that need it skip when it is absent.
"""
import ctypes as C
import os
import numpy as np

from vpic_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ORACLE_DIR = os.path.join(ROOT, "oracle")

_ref_libs = {}
_booted = set()


def ref_path(variant="scalar"):
    return os.path.join(ORACLE_DIR, "_ref", f"libvpic_ref_{variant}.so")


def have_ref(variant="scalar"):
    return os.path.exists(ref_path(variant))


def load_oracle():
    """liboracle.so — the plain-C port."""
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])
    lib = C.CDLL(path)
    lib.vpo_advance_p.restype = C.c_int32
    lib.vpo_energy_p.restype = C.c_double
    lib.vpo_energy_p.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float]
    lib.vpo_unload_accumulator.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_float, C.c_float, C.c_float, C.c_float]
    lib.vpo_center_p.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float]
    lib.vpo_uncenter_p.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float]
    lib.vpo_advance_p.argtypes = [C.c_void_p, C.c_void_p]
    lib.vpo_move_p.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_float]
    lib.vpo_sort_p.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.vpo_load_interpolator.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.vpo_clear_accumulator.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.vpo_clear_jf.argtypes = [C.c_void_p]
    lib.vpo_synchronize_jf.argtypes = [C.c_void_p]
    lib.vpo_vacuum_energy_f.argtypes = [C.c_void_p, C.c_void_p]
    lib.vpo_accumulate_hydro_p.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32] + [C.c_float] * 5 + [C.c_int32] * 3
    lib.vpo_accumulate_hydro_p.restype = None
    lib.vpo_synchronize_hydro.argtypes = [C.c_void_p, C.c_void_p]
    lib.vpo_synchronize_hydro.restype = None
    for name in ("vpo_clear_rhof", "vpo_synchronize_rho", "vpo_vacuum_compute_div_e_err", "vpo_vacuum_clean_div_e",
                 "vpo_compute_div_b_err", "vpo_clean_div_b", "vpo_vacuum_compute_rhob", "vpo_vacuum_compute_curl_b"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = None
    for name in ("vpo_compute_rms_div_e_err", "vpo_compute_rms_div_b_err", "vpo_synchronize_tang_e_norm_b"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = C.c_double
    lib.vpo_accumulate_rho_p.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_int32]
    lib.vpo_accumulate_rhob.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_int32]
    lib.vpo_advance_b.argtypes = [C.c_void_p, C.c_float]
    lib.vpo_vacuum_advance_e.argtypes = [C.c_void_p, C.c_float]
    return lib


class OraclePushArgs(C.Structure):
    _fields_ = [("p", C.c_void_p), ("np", C.c_int32),
                ("pm", C.c_void_p), ("max_nm", C.c_int32),
                ("interp", C.c_void_p), ("interp_stride", C.c_int32),
                ("accum", C.c_void_p), ("accum_stride", C.c_int32),
                ("neighbor", C.c_void_p),
                ("rangel", C.c_int64), ("rangeh", C.c_int64),
                ("qdt_2mc", C.c_float), ("cdt_dx", C.c_float), ("cdt_dy", C.c_float),
                ("cdt_dz", C.c_float), ("qsp", C.c_float)]


class OracleFieldArgs(C.Structure):
    _fields_ = [("f", C.c_void_p), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("dt", C.c_float), ("cvac", C.c_float), ("eps0", C.c_float), ("damp", C.c_float),
                ("dx", C.c_float), ("dy", C.c_float), ("dz", C.c_float), ("dV", C.c_float),
                ("rdx", C.c_float), ("rdy", C.c_float), ("rdz", C.c_float),
                ("bc6", C.c_int32 * 6), ("has_material", C.c_int32), ("material", C.c_float * 13)]


def load_ref(variant="scalar", tpp=1):
    """Load and boot one reference build.  One boot per process per variant."""
    if variant in _ref_libs:
        return _ref_libs[variant]
    lib = C.CDLL(ref_path(variant), mode=os.RTLD_LAZY | os.RTLD_LOCAL)  # deck hooks (user_*) stay unresolved
    args = [b"refvpic", b"--tpp", str(tpp).encode()]
    argc = C.c_int(len(args))
    argv_arr = (C.c_char_p * (len(args) + 1))(*args, None)
    argv = C.cast(argv_arr, C.POINTER(C.c_char_p))
    pargv = C.pointer(argv)
    lib.boot_services(C.byref(argc), pargv)
    lib._keep = (argv_arr, argv, pargv)

    lib.new_grid.restype = C.POINTER(abi.Grid)
    lib.partition_periodic_box.argtypes = [C.POINTER(abi.Grid)] + [C.c_double] * 6 + [C.c_int] * 6
    lib.partition_metal_box.argtypes = [C.POINTER(abi.Grid)] + [C.c_double] * 6 + [C.c_int] * 6
    lib.set_fbc.argtypes = [C.POINTER(abi.Grid), C.c_int, C.c_int]
    lib.set_pbc.argtypes = [C.POINTER(abi.Grid), C.c_int, C.c_int]
    lib.material.restype = C.c_void_p
    lib.material.argtypes = [C.c_char_p] + [C.c_float] * 12
    lib.append_material.restype = C.c_void_p
    lib.append_material.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.new_standard_field_array.restype = C.POINTER(abi.FieldArray)
    lib.new_standard_field_array.argtypes = [C.POINTER(abi.Grid), C.c_void_p, C.c_float]
    lib.new_interpolator_array.restype = C.POINTER(abi.InterpolatorArray)
    lib.new_interpolator_array.argtypes = [C.POINTER(abi.Grid)]
    lib.new_accumulator_array.restype = C.POINTER(abi.AccumulatorArray)
    lib.new_accumulator_array.argtypes = [C.POINTER(abi.Grid)]
    lib.species.restype = C.POINTER(abi.Species)
    lib.species.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                            C.POINTER(abi.Grid)]
    for fn in ("advance_p",):
        getattr(lib, fn).argtypes = [C.POINTER(abi.Species), C.POINTER(abi.AccumulatorArray),
                                     C.POINTER(abi.InterpolatorArray)]
    lib.sort_p.argtypes = [C.POINTER(abi.Species)]
    lib.center_p.argtypes = [C.POINTER(abi.Species), C.POINTER(abi.InterpolatorArray)]
    lib.uncenter_p.argtypes = [C.POINTER(abi.Species), C.POINTER(abi.InterpolatorArray)]
    lib.energy_p.argtypes = [C.POINTER(abi.Species), C.POINTER(abi.InterpolatorArray)]
    lib.energy_p.restype = C.c_double
    lib.load_interpolator_array.argtypes = [C.POINTER(abi.InterpolatorArray), C.POINTER(abi.FieldArray)]
    lib.clear_accumulator_array.argtypes = [C.POINTER(abi.AccumulatorArray)]
    lib.reduce_accumulator_array.argtypes = [C.POINTER(abi.AccumulatorArray)]
    lib.unload_accumulator_array.argtypes = [C.POINTER(abi.FieldArray), C.POINTER(abi.AccumulatorArray)]
    lib.accumulate_rho_p.argtypes = [C.POINTER(abi.FieldArray), C.POINTER(abi.Species)]
    lib.accumulate_rhob.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(abi.Grid), C.c_float]
    lib.new_hydro_array.restype = C.POINTER(abi.HydroArray)
    lib.new_hydro_array.argtypes = [C.POINTER(abi.Grid)]
    lib.clear_hydro_array.argtypes = [C.POINTER(abi.HydroArray)]
    lib.synchronize_hydro_array.argtypes = [C.POINTER(abi.HydroArray)]
    lib.accumulate_hydro_p.argtypes = [C.POINTER(abi.HydroArray), C.POINTER(abi.Species), C.POINTER(abi.InterpolatorArray)]
    lib.move_p.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(abi.Grid), C.c_float]
    lib.move_p.restype = C.c_int
    lib.variant = variant
    lib.simd_width = {"scalar": 4, "v4": 4, "v8": 8, "v16": 16}[variant]
    _ref_libs[variant] = lib
    return lib


def BOUNDARY(i, j, k):
    return 13 + i + 3 * j + 9 * k


FACES = [(-1, 0, 0), (0, -1, 0), (0, 0, -1), (1, 0, 0), (0, 1, 0), (0, 0, 1)]


class RefWorld:
    """A reference grid + field/interpolator/accumulator arrays, with numpy views of the host arrays."""

    def __init__(self, lib, nx, ny, nz, lx=None, ly=None, lz=None, dt=None, cvac=1.0, eps0=1.0, damp=0.0,
                 fbc=None, pbc=None, material=None):
        """material: 12 floats eps(x,y,z), mu(x,y,z), sigma(x,y,z), zeta(x,y,z) of the single material that fills
        space (material.h); default vacuum."""
        self.lib = lib
        g = lib.new_grid()
        self.g = g
        lx = lx if lx is not None else float(nx)
        ly = ly if ly is not None else float(ny)
        lz = lz if lz is not None else float(nz)
        g.contents.cvac = cvac
        g.contents.eps0 = eps0
        lib.partition_periodic_box(g, 0., 0., 0., lx, ly, lz, nx, ny, nz, 1, 1, 1)
        if dt is None:
            inv = sum((1.0 / d) ** 2 for d, n in ((lx / nx, nx), (ly / ny, ny), (lz / nz, nz)) if n > 1)
            dt = 0.98 / (cvac * np.sqrt(inv))
        g.contents.dt = dt
        # optional local boundary conditions per face index 0..5 (-x,-y,-z,+x,+y,+z)
        for f, code in (fbc or {}).items():
            lib.set_fbc(g, BOUNDARY(*FACES[f]), code)
        for f, code in (pbc or {}).items():
            lib.set_pbc(g, BOUNDARY(*FACES[f]), code)
        m_list = C.c_void_p(None)
        m = lib.material(b"vacuum", *(list(material) if material is not None else [1.0] * 6 + [0.0] * 6))
        lib.append_material(m, C.byref(m_list))
        self.fa = lib.new_standard_field_array(g, m_list, damp)
        self.ia = lib.new_interpolator_array(g)
        self.aa = lib.new_accumulator_array(g)
        self.nx, self.ny, self.nz, self.nv = nx, ny, nz, g.contents.nv
        self.damp = damp
        self.isf = abi.interpolator_floats(lib.simd_width)
        self.asf = abi.accumulator_floats(lib.simd_width)

    # numpy views over host memory owned by the reference
    @property
    def fields(self):
        return np.ctypeslib.as_array(self.fa.contents.f, shape=(self.nv, abi.FIELD_FLOATS))

    @property
    def interp(self):
        return np.ctypeslib.as_array(self.ia.contents.i, shape=(self.nv, self.isf))

    @property
    def accum(self):
        aa = self.aa.contents
        return np.ctypeslib.as_array(aa.a, shape=(aa.n_pipeline + 1, aa.stride, self.asf))

    @property
    def neighbor(self):
        return np.ctypeslib.as_array(self.g.contents.neighbor, shape=(self.nv, 6))

    def kernel(self, idx, restype, *argtypes):
        """Function pointer idx of fa->kernel[0] (field_advance.h:170-218)."""
        return C.CFUNCTYPE(restype, *argtypes)(self.fa.contents.kernel[idx])

    def advance_b(self, frac):
        self.kernel(1, None, C.POINTER(abi.FieldArray), C.c_float)(self.fa, frac)

    def advance_e(self, frac=1.0):
        self.kernel(2, None, C.POINTER(abi.FieldArray), C.c_float)(self.fa, frac)

    def energy_f(self):
        en = (C.c_double * 6)()
        self.kernel(3, None, C.POINTER(C.c_double), C.POINTER(abi.FieldArray))(en, self.fa)
        return np.array(en[:])

    def clear_jf(self):
        self.kernel(4, None, C.POINTER(abi.FieldArray))(self.fa)

    def synchronize_jf(self):
        self.kernel(5, None, C.POINTER(abi.FieldArray))(self.fa)

    # divergence cleaning / shared-face synchronisation entries (field_advance.h:186-218)
    def _k_void(self, idx):
        self.kernel(idx, None, C.POINTER(abi.FieldArray))(self.fa)

    def _k_double(self, idx):
        return float(self.kernel(idx, C.c_double, C.POINTER(abi.FieldArray))(self.fa))

    def compute_rhob(self): self._k_void(8)
    def compute_curl_b(self): self._k_void(9)
    def clear_rhof(self): self._k_void(6)
    def synchronize_rho(self): self._k_void(7)
    def synchronize_tang_e_norm_b(self): return self._k_double(10)
    def compute_div_e_err(self): self._k_void(11)
    def compute_rms_div_e_err(self): return self._k_double(12)
    def clean_div_e(self): self._k_void(13)
    def compute_div_b_err(self): self._k_void(14)
    def compute_rms_div_b_err(self): return self._k_double(15)
    def clean_div_b(self): self._k_void(16)

    def new_species(self, name, q, m, max_np, max_nm, sort_interval=20):
        sp = self.lib.species(name.encode(), q, m, max_np, max_nm, sort_interval, 0, self.g)
        return RefSpecies(self, sp)

    def field_args(self, farr):
        """vpo_field_args_t for the port, sharing this world's grid constants."""
        g = self.g.contents
        a = OracleFieldArgs()
        a.f = farr.ctypes.data
        a.nx, a.ny, a.nz = self.nx, self.ny, self.nz
        a.dt, a.cvac, a.eps0, a.damp = g.dt, g.cvac, g.eps0, self.damp
        a.dx, a.dy, a.dz, a.dV = g.dx, g.dy, g.dz, g.dV
        a.rdx, a.rdy, a.rdz = g.rdx, g.rdy, g.rdz
        for f in range(6):
            a.bc6[f] = g.bc[BOUNDARY(*FACES[f])]
        a.has_material = 1
        for k, v in enumerate(self.material_coefficients()):
            a.material[k] = v
        return a

    def material_coefficients(self):
        """The 13 material_coefficient_t floats the reference derived for the single material (sfa.cc:108-148)."""
        prm = C.cast(self.fa.contents.params, C.POINTER(abi.SfaParams)).contents
        mc = C.cast(prm.mc, C.POINTER(C.c_float * 13)).contents
        return [float(x) for x in mc]


class RefSpecies:
    def __init__(self, world, sp):
        self.world, self.sp = world, sp

    @property
    def c(self):
        return self.sp.contents

    @property
    def p(self):
        return np.ctypeslib.as_array(C.cast(self.c.p, C.POINTER(C.c_byte)),
                                     shape=(self.c.max_np * 32,)).view(abi.particle_dtype)

    @property
    def pm(self):
        return np.ctypeslib.as_array(C.cast(self.c.pm, C.POINTER(C.c_byte)),
                                     shape=(self.c.max_nm * 16,)).view(abi.mover_dtype)

    @property
    def partition(self):
        return np.ctypeslib.as_array(self.c.partition, shape=(self.world.nv + 1,))

    def set_particles(self, arr):
        n = len(arr)
        assert n <= self.c.max_np
        self.p[:n] = arr
        self.c.np = n
        self.c.nm = 0

    def push_constants(self):
        """The float constants advance_p_pipeline computes on the host (advance_p_pipeline.cc:279-283)."""
        g = self.world.g.contents
        f32 = np.float32
        q, m, dt, cvac = f32(self.c.q), f32(self.c.m), f32(g.dt), f32(g.cvac)
        qdt_2mc = f32(f32(q * dt) / f32(f32(f32(2) * m) * cvac))
        return dict(qdt_2mc=qdt_2mc,
                    cdt_dx=f32(f32(cvac * dt) * f32(g.rdx)),
                    cdt_dy=f32(f32(cvac * dt) * f32(g.rdy)),
                    cdt_dz=f32(f32(cvac * dt) * f32(g.rdz)),
                    qsp=q)


def random_particles(rng, n, nx, ny, nz, uth=0.2, w=1.0, drift=(0, 0, 0)):
    """n particles uniformly placed in interior voxels with thermal momenta."""
    p = np.zeros(n, dtype=abi.particle_dtype)
    p["dx"] = rng.uniform(-1, 1, n).astype(np.float32)
    p["dy"] = rng.uniform(-1, 1, n).astype(np.float32)
    p["dz"] = rng.uniform(-1, 1, n).astype(np.float32)
    ix = rng.integers(1, nx + 1, n)
    iy = rng.integers(1, ny + 1, n)
    iz = rng.integers(1, nz + 1, n)
    p["i"] = abi.voxel(ix, iy, iz, nx, ny, nz).astype(np.int32)
    for k, d in zip(("ux", "uy", "uz"), drift):
        p[k] = (rng.normal(0, uth, n) + d).astype(np.float32)
    p["w"] = np.float32(w)
    return p


def random_fields(rng, nv, amp_e=0.05, amp_b=0.05):
    f = np.zeros((nv, abi.FIELD_FLOATS), dtype=np.float32)
    f[:, 0:3] = rng.normal(0, amp_e, (nv, 3))
    f[:, 4:7] = rng.normal(0, amp_b, (nv, 3))
    return f

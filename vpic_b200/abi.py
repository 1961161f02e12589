"""ctypes mirror of include/vpic_b200_abi.h — the reference's struct layouts.

These are the binary layouts the drop-in boundary honours
(species_advance_aos.h:21-94, grid.h:73-131, sf_interface.h:62-131,
field_advance.h:152-229 in the reference).  `simd_width` selects the
interpolator/accumulator padding of the host build (sf_interface.h:27-53).
"""
import ctypes as C
import numpy as np

c_f, c_i32, c_i64 = C.c_float, C.c_int32, C.c_int64


class Particle(C.Structure):
    _fields_ = [("dx", c_f), ("dy", c_f), ("dz", c_f), ("i", c_i32),
                ("ux", c_f), ("uy", c_f), ("uz", c_f), ("w", c_f)]


class ParticleMover(C.Structure):
    _fields_ = [("dispx", c_f), ("dispy", c_f), ("dispz", c_f), ("i", c_i32)]


class ParticleInjector(C.Structure):
    _fields_ = [("dx", c_f), ("dy", c_f), ("dz", c_f), ("i", c_i32),
                ("ux", c_f), ("uy", c_f), ("uz", c_f), ("w", c_f),
                ("dispx", c_f), ("dispy", c_f), ("dispz", c_f), ("sp_id", c_i32)]


class Grid(C.Structure):
    _fields_ = [("dt", c_f), ("cvac", c_f), ("eps0", c_f),
                ("step", c_i64), ("t0", C.c_double),
                ("x0", c_f), ("y0", c_f), ("z0", c_f), ("x1", c_f), ("y1", c_f), ("z1", c_f),
                ("nx", c_i32), ("ny", c_i32), ("nz", c_i32),
                ("dx", c_f), ("dy", c_f), ("dz", c_f), ("dV", c_f),
                ("rdx", c_f), ("rdy", c_f), ("rdz", c_f), ("r8V", c_f),
                ("sx", c_i32), ("sy", c_i32), ("sz", c_i32), ("nv", c_i32),
                ("bc", c_i32 * 27),
                ("range", C.POINTER(c_i64)),
                ("neighbor", C.POINTER(c_i64)),
                ("rangel", c_i64), ("rangeh", c_i64),
                ("mp", C.c_void_p)]


class Species(C.Structure):
    pass


Species._fields_ = [("name", C.c_char_p), ("q", c_f), ("m", c_f),
                    ("np", c_i32), ("max_np", c_i32), ("p", C.POINTER(Particle)),
                    ("nm", c_i32), ("max_nm", c_i32), ("pm", C.POINTER(ParticleMover)),
                    ("last_sorted", c_i64), ("sort_interval", c_i32), ("sort_out_of_place", c_i32),
                    ("partition", C.POINTER(c_i32)),
                    ("g", C.POINTER(Grid)), ("id", c_i32), ("next", C.POINTER(Species))]


def interpolator_floats(simd_width=4):
    return {4: 20, 8: 24, 16: 32}[simd_width]


def accumulator_floats(simd_width=4):
    return {4: 12, 8: 16, 16: 16}[simd_width]


class InterpolatorArray(C.Structure):
    _fields_ = [("i", C.POINTER(c_f)), ("g", C.POINTER(Grid))]


class AccumulatorArray(C.Structure):
    _fields_ = [("a", C.POINTER(c_f)), ("n_pipeline", c_i32), ("stride", c_i32), ("g", C.POINTER(Grid))]


class HydroArray(C.Structure):
    """hydro_array_t (sf_interface.h:194-200); hydro_t is 16 floats (14 moments + 2 pad) for every SIMD width."""
    _fields_ = [("h", C.POINTER(c_f)), ("n_pipeline", c_i32), ("stride", c_i32), ("g", C.POINTER(Grid))]


HYDRO_FLOATS = 16
FIELD_FLOATS = 20  # 80-byte field_t


class FieldArray(C.Structure):
    _fields_ = [("f", C.POINTER(c_f)), ("g", C.POINTER(Grid)), ("params", C.c_void_p),
                ("kernel", C.c_void_p * 17)]


class MaterialCoefficient(C.Structure):
    _fields_ = [(n, c_f) for n in ("decayx", "drivex", "decayy", "drivey", "decayz", "drivez",
                                   "rmux", "rmuy", "rmuz", "nonconductive", "epsx", "epsy", "epsz")] + \
               [("pad_", c_f * 3)]


class SfaParams(C.Structure):
    _fields_ = [("mc", C.POINTER(MaterialCoefficient)), ("n_mc", c_i32), ("damp", c_f)]


# numpy views of the same layouts
particle_dtype = np.dtype([("dx", "f4"), ("dy", "f4"), ("dz", "f4"), ("i", "i4"),
                           ("ux", "f4"), ("uy", "f4"), ("uz", "f4"), ("w", "f4")])
mover_dtype = np.dtype([("dispx", "f4"), ("dispy", "f4"), ("dispz", "f4"), ("i", "i4")])
injector_dtype = np.dtype([("dx", "f4"), ("dy", "f4"), ("dz", "f4"), ("i", "i4"),
                           ("ux", "f4"), ("uy", "f4"), ("uz", "f4"), ("w", "f4"),
                           ("dispx", "f4"), ("dispy", "f4"), ("dispz", "f4"), ("sp_id", "i4")])

# field_t member offsets (floats)
F = dict(ex=0, ey=1, ez=2, div_e_err=3, cbx=4, cby=5, cbz=6, div_b_err=7,
         tcax=8, tcay=9, tcaz=10, rhob=11, jfx=12, jfy=13, jfz=14, rhof=15)
# interpolator_t member offsets (floats)
I = dict(ex=0, dexdy=1, dexdz=2, d2exdydz=3, ey=4, deydz=5, deydx=6, d2eydzdx=7,
         ez=8, dezdx=9, dezdy=10, d2ezdxdy=11, cbx=12, dcbxdx=13, cby=14, dcbydy=15, cbz=16, dcbzdz=17)


def voxel(x, y, z, nx, ny, nz):
    """VOXEL macro, grid.h:136."""
    return x + (nx + 2) * (y + (ny + 2) * z)

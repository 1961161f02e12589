"""ctypes binding of libvpic_b200.so (include/vpic_b200.h).

There is no fallback: if the CUDA library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvpic_b200.so")

c_f, c_i32, c_i64, c_vp = C.c_float, C.c_int32, C.c_int64, C.c_void_p


class VpbError(RuntimeError):
    pass


class NeighborRule(C.Structure):
    """vpb_neighbor_rule_t"""
    _fields_ = [("valid", c_i32), ("nx", c_i32), ("ny", c_i32), ("nz", c_i32), ("act", c_i64 * 6), ("delta", c_i64 * 6)]


class PushArgs(C.Structure):
    """vpb_push_args_t"""
    _fields_ = [("p", c_vp), ("np", c_i32),
                ("pm", c_vp), ("max_nm", c_i32),
                ("counters", c_vp),
                ("interp", c_vp), ("interp_stride", c_i32),
                ("accum", c_vp), ("accum_stride", c_i32),
                ("neighbor", c_vp), ("rangel", c_i64), ("rangeh", c_i64),
                ("qdt_2mc", c_f), ("cdt_dx", c_f), ("cdt_dy", c_f), ("cdt_dz", c_f), ("qsp", c_f),
                ("nx", c_i32), ("ny", c_i32), ("nz", c_i32),
                ("variant", c_i32), ("p_first", c_i32), ("neighbor_rule", C.POINTER(NeighborRule)), ("debug_skip", c_i32),
                ("partition", c_vp), ("partition_np", c_i32),
                ("perm", c_vp), ("p_out", c_vp), ("keys_out", c_vp)]


class BoundaryArgs(C.Structure):
    """vpb_boundary_args_t"""
    _fields_ = [("p", c_vp), ("np", c_i32), ("pm", c_vp), ("nm", c_i32), ("neighbor", c_vp),
                ("rangel", c_i64), ("rangeh", c_i64), ("rangem", c_i64), ("face_range", c_i64 * 6),
                ("sp_id", c_i32), ("inj", c_vp), ("class_offsets", c_vp), ("scratch", c_vp), ("scratch_bytes", C.c_size_t),
                ("fields", c_vp), ("q_r8V", c_f), ("nx", c_i32), ("ny", c_i32), ("nz", c_i32), ("absorb_all", c_i32)]


class FieldArgs(C.Structure):
    """vpb_field_args_t"""
    _fields_ = [("f", c_vp), ("nx", c_i32), ("ny", c_i32), ("nz", c_i32),
                ("dt", c_f), ("cvac", c_f), ("eps0", c_f), ("damp", c_f),
                ("dx", c_f), ("dy", c_f), ("dz", c_f), ("dV", c_f),
                ("rdx", c_f), ("rdy", c_f), ("rdz", c_f),
                ("face", c_i32 * 6), ("has_material", c_i32), ("material", c_f * 13)]


DEPOSIT_DEFAULT, DEPOSIT_RED_V4, DEPOSIT_WARP_SEG, DEPOSIT_WARP_SEG_MOVERS, DEPOSIT_WARP_SEG_FIRST = 0, 1, 2, 3, 4
DEPOSIT_BRICK_TILE = 5
FACE_PERIODIC_SELF, FACE_REMOTE = 0, 1
HALO_TANG_B, HALO_JF, HALO_RHO, HALO_NORM_E, HALO_DIV_B, HALO_TANG_E_NORM_B = 0, 1, 2, 3, 4, 5

# every exported symbol of include/vpic_b200.h: name -> (restype, argtypes)
_PROTOS = {
    "vpb_version": (C.c_int, []),
    "vpb_last_error": (C.c_char_p, []),
    "vpb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "vpb_set_device": (C.c_int, [C.c_int]),
    "vpb_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "vpb_malloc": (C.c_int, [C.POINTER(c_vp), C.c_size_t]),
    "vpb_free": (C.c_int, [c_vp]),
    "vpb_malloc_host": (C.c_int, [C.POINTER(c_vp), C.c_size_t]),
    "vpb_free_host": (C.c_int, [c_vp]),
    "vpb_memset": (C.c_int, [c_vp, C.c_int, C.c_size_t, c_vp]),
    "vpb_memcpy_h2d": (C.c_int, [c_vp, c_vp, C.c_size_t, c_vp]),
    "vpb_memcpy_d2h": (C.c_int, [c_vp, c_vp, C.c_size_t, c_vp]),
    "vpb_memcpy_d2d": (C.c_int, [c_vp, c_vp, C.c_size_t, c_vp]),
    "vpb_stream_sync": (C.c_int, [c_vp]),
    "vpb_device_sync": (C.c_int, []),
    "vpb_launch_count": (c_i64, []),
    "vpb_advance_p": (C.c_int, [C.POINTER(PushArgs), c_vp]),
    "vpb_neighbor_rule_derive": (C.c_int, [c_vp, c_i32, c_i32, c_i32, c_i64, C.POINTER(NeighborRule), c_vp]),
    "vpb_boundary_scratch_bytes": (C.c_size_t, [c_i32]),
    "vpb_boundary_p_pack": (C.c_int, [C.POINTER(BoundaryArgs), c_vp]),
    "vpb_boundary_p_inject": (C.c_int, [C.POINTER(PushArgs), c_vp, c_i32, c_vp]),
    "vpb_move_p": (C.c_int, [C.POINTER(PushArgs), c_vp, c_vp, c_vp]),
    "vpb_boundary_msg_bytes": (C.c_size_t, [c_i32]),
    "vpb_boundary_p_stage": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "vpb_boundary_p_inject_msg": (C.c_int, [C.POINTER(PushArgs), c_vp, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "vpb_sort_scratch_bytes": (C.c_size_t, [c_i32, c_i32]),
    "vpb_sort_movers_scratch_bytes": (C.c_size_t, [c_i32]),
    "vpb_sort_movers": (C.c_int, [c_vp, c_i32, c_vp, C.c_size_t, c_vp]),
    "vpb_sort_p": (C.c_int, [c_vp, c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, C.c_size_t, c_vp]),
    "vpb_sort_index_work_bytes": (C.c_size_t, [c_i32]),
    "vpb_sort_index_scratch_bytes": (C.c_size_t, [c_i32, c_i32]),
    "vpb_sort_p_index": (C.c_int, [c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, C.c_size_t, c_vp, C.c_size_t, c_vp, c_vp]),
    "vpb_permute_p": (C.c_int, [c_vp, c_i32, c_vp, c_vp, c_vp]),
    "vpb_unpermute_p": (C.c_int, [c_vp, c_i32, c_vp, c_vp, c_vp]),
    "vpb_extract_keys": (C.c_int, [c_vp, c_i32, c_vp, c_vp]),
    "vpb_load_interpolator": (C.c_int, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "vpb_clear_accumulator": (C.c_int, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "vpb_unload_accumulator": (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f, c_f, c_f, c_f, c_vp]),
    "vpb_accumulate_rho_p": (C.c_int, [c_vp, c_vp, c_i32, c_f, c_f, c_i32, c_i32, c_i32, c_vp]),
    "vpb_energy_p": (C.c_int, [c_vp, c_i32, c_vp, c_i32, c_f, c_f, c_f, c_f, c_vp, c_vp]),
    "vpb_center_p": (C.c_int, [c_vp, c_i32, c_vp, c_i32, c_f, c_vp]),
    "vpb_uncenter_p": (C.c_int, [c_vp, c_i32, c_vp, c_i32, c_f, c_vp]),
    "vpb_advance_b": (C.c_int, [C.POINTER(FieldArgs), c_f, c_vp]),
    "vpb_vacuum_advance_e": (C.c_int, [C.POINTER(FieldArgs), c_f, c_vp]),
    "vpb_clear_jf": (C.c_int, [C.POINTER(FieldArgs), c_vp]),
    "vpb_synchronize_jf": (C.c_int, [C.POINTER(FieldArgs), c_vp]),
    "vpb_vacuum_energy_f": (C.c_int, [C.POINTER(FieldArgs), c_vp, c_vp]),
    "vpb_clear_rhof": (C.c_int, [C.POINTER(FieldArgs), c_vp]),
    "vpb_synchronize_rho": (C.c_int, [C.POINTER(FieldArgs), c_vp]),
    "vpb_vacuum_compute_div_e_err": (C.c_int, [C.POINTER(FieldArgs), c_vp]),
    "vpb_compute_rms_div_e_err": (C.c_int, [C.POINTER(FieldArgs), c_vp, c_vp]),
    "vpb_vacuum_clean_div_e": (C.c_int, [C.POINTER(FieldArgs), c_vp]),
    "vpb_compute_div_b_err": (C.c_int, [C.POINTER(FieldArgs), c_vp]),
    "vpb_compute_rms_div_b_err": (C.c_int, [C.POINTER(FieldArgs), c_vp, c_vp]),
    "vpb_clean_div_b": (C.c_int, [C.POINTER(FieldArgs), c_vp]),
    "vpb_synchronize_tang_e_norm_b": (C.c_int, [C.POINTER(FieldArgs), c_vp, c_vp]),
    "vpb_vacuum_compute_rhob": (C.c_int, [C.POINTER(FieldArgs), c_vp]),
    "vpb_vacuum_compute_curl_b": (C.c_int, [C.POINTER(FieldArgs), c_vp]),
    "vpb_accumulate_hydro_p": (C.c_int, [c_vp, c_vp, c_i32, c_vp, c_i32] + [c_f] * 5 + [c_i32] * 3 + [c_vp]),
    "vpb_clear_hydro": (C.c_int, [c_vp, c_i32, c_i32, c_i32, c_vp]),
    "vpb_synchronize_hydro": (C.c_int, [c_vp, C.POINTER(FieldArgs), c_vp]),
    "vpb_hydro_halo_floats": (C.c_size_t, [c_i32, c_i32, c_i32, C.c_int]),
    "vpb_hydro_halo_pack": (C.c_int, [c_vp, C.POINTER(FieldArgs), C.c_int, c_vp, c_vp]),
    "vpb_hydro_halo_unpack": (C.c_int, [c_vp, C.POINTER(FieldArgs), C.c_int, c_vp, c_vp]),
    "vpb_halo_floats": (C.c_size_t, [c_i32, c_i32, c_i32, C.c_int]),
    "vpb_halo_floats_kind": (C.c_size_t, [c_i32, c_i32, c_i32, C.c_int, C.c_int]),
    "vpb_halo_unpack_sync": (C.c_int, [C.POINTER(FieldArgs), C.c_int, c_vp, c_vp, c_vp]),
    "vpb_halo_pack": (C.c_int, [C.POINTER(FieldArgs), C.c_int, C.c_int, c_vp, c_vp]),
    "vpb_halo_unpack": (C.c_int, [C.POINTER(FieldArgs), C.c_int, C.c_int, c_vp, c_vp]),
}

# the reference's own extern "C" symbols that the drop-in layer exports (include/vpic_b200_dropin.h)
DROPIN_SYMBOLS = ["advance_p", "sort_p", "move_p", "load_interpolator_array", "clear_accumulator_array",
                  "reduce_accumulator_array", "unload_accumulator_array", "energy_p", "center_p", "uncenter_p",
                  "accumulate_rho_p", "advance_b", "vacuum_advance_e", "clear_jf", "synchronize_jf", "vacuum_energy_f",
                  "clear_rhof", "synchronize_rho", "vacuum_compute_div_e_err", "compute_rms_div_e_err", "vacuum_clean_div_e",
                  "compute_div_b_err", "compute_rms_div_b_err", "clean_div_b", "synchronize_tang_e_norm_b",
                  "vacuum_compute_rhob", "vacuum_compute_curl_b", "vpic_b200_compute_rhob", "vpic_b200_compute_curl_b",
                  "boundary_p", "accumulate_hydro_p", "clear_hydro_array", "reduce_hydro_array", "synchronize_hydro_array",
                  "vpic_b200_clear_rhof", "vpic_b200_synchronize_rho", "vpic_b200_compute_div_e_err",
                  "vpic_b200_compute_rms_div_e_err", "vpic_b200_clean_div_e", "vpic_b200_compute_div_b_err",
                  "vpic_b200_compute_rms_div_b_err", "vpic_b200_clean_div_b", "vpic_b200_synchronize_tang_e_norm_b",
                  "vpic_b200_advance_b", "vpic_b200_advance_e", "vpic_b200_clear_jf", "vpic_b200_synchronize_jf",
                  "vpic_b200_energy_f", "vpic_b200_install_field_kernels",
                  "vpic_b200_sync_to_host", "vpic_b200_invalidate", "vpic_b200_release", "vpic_b200_set_mode",
                  "vpic_b200_transfer_bytes", "vpic_b200_host_access", "vpic_b200_lazy_stats", "vpic_b200_set_lazy_min"]

_lib = None


def exported_symbols():
    return list(_PROTOS) + DROPIN_SYMBOLS


def load():
    """Load libvpic_b200.so (build it first: python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VpbError(f"{LIB_PATH} is missing — build the CUDA extension (make -C vpic_b200/csrc); "
                       "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().vpb_last_error().decode(errors="replace")
        raise VpbError(f"{what} failed (rc={rc}): {msg}")

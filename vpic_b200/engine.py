"""Host-side mirror of the reference's operator interface for the hot path, over device-resident arrays.

Names, argument meaning and error behaviour follow the reference's C API
(src/species_advance/species_advance.h:27-159, src/sf_interface/sf_interface.h:82-174): `advance_p(sp, aa, ia)`,
`sort_p(sp)`, `load_interpolator_array(ia, fa)`, `clear/reduce/unload_accumulator_array`, `energy_p`, ...
Objects hold torch CUDA tensors in the reference's exact layouts (particle_t 32 B, interpolator_t 80 B,
accumulator_t 48 B, field_t 80 B); every operator is one call into libvpic_b200.so through the C-ABI.
torch is plumbing only (device memory, streams, torch.distributed) — there is no torch math on this path and no
CPU fallback: without the CUDA library these calls raise.
"""
import ctypes as C
import os
import numpy as np
import torch

from . import abi, lib as _lib
from .grid import Grid


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


def _bad_args(cond, who):
    if cond:
        raise ValueError(f"{who}: Bad args.")   # the reference ERROR()s and exits; a library raises instead


class DeviceGrid:
    """Device copy of the read-only parts of grid_t."""

    def __init__(self, g: Grid, device="cuda"):
        self.g = g
        self.device = torch.device(device)
        self.neighbor = torch.from_numpy(np.ascontiguousarray(g.neighbor)).to(self.device)
        for k in ("nx", "ny", "nz", "nv", "dt", "cvac", "eps0", "dx", "dy", "dz", "dV", "rdx", "rdy", "rdz",
                  "rangel", "rangeh", "rank", "world_size"):
            setattr(self, k, getattr(g, k))

        self.use_neighbor_rule = os.environ.get("VPB_NO_NEIGHBOR_RULE", "0") != "1"
        self._rule = None

    @property
    def step(self):
        return self.g.step

    def neighbor_rule(self):
        """Verified closed form of the neighbour table (vpb_neighbor_rule_derive), or None to use the table."""
        if not self.use_neighbor_rule:
            return None
        if self._rule is None:
            r = _lib.NeighborRule()
            _lib.check(_lib.load().vpb_neighbor_rule_derive(_ptr(self.neighbor), self.nx, self.ny, self.nz, self.rangel,
                                                            C.byref(r), _stream()), "neighbor_rule_derive")
            self._rule = r
        return C.pointer(self._rule) if self._rule.valid else None


class InterpolatorArray:
    def __init__(self, g: DeviceGrid, simd_width=4):
        self.g = g
        self.stride = abi.interpolator_floats(simd_width)
        self.i = torch.zeros((g.nv, self.stride), dtype=torch.float32, device=g.device)


class AccumulatorArray:
    """One accumulator block on the device (reduce_accumulator_array is the identity)."""

    def __init__(self, g: DeviceGrid, simd_width=4):
        self.g = g
        self.stride_floats = abi.accumulator_floats(simd_width)
        self.stride = (g.nv + 1) // 2 * 2                      # POW2_CEIL(nv,2), accumulator_array.cc:78
        self.n_pipeline = 0
        self.a = torch.zeros((self.stride, self.stride_floats), dtype=torch.float32, device=g.device)


class FieldArray:
    def __init__(self, g: DeviceGrid, damp=0.0, material=None):
        """material: the 13 material_coefficient_t floats (sfa_private.h:14-25) of the single material that fills
        space, or None for true vacuum."""
        self.g = g
        self.damp = float(damp)
        self.material = None if material is None else [float(x) for x in material]
        assert self.material is None or len(self.material) == 13
        self.f = torch.zeros((g.nv, abi.FIELD_FLOATS), dtype=torch.float32, device=g.device)
        self._en = torch.zeros(6, dtype=torch.float64, device=g.device)

    def args(self):
        g = self.g
        a = _lib.FieldArgs()
        a.f = self.f.data_ptr()
        a.nx, a.ny, a.nz = g.nx, g.ny, g.nz
        a.dt, a.cvac, a.eps0, a.damp = g.dt, g.cvac, g.eps0, self.damp
        a.dx, a.dy, a.dz, a.dV = g.dx, g.dy, g.dz, g.dV
        a.rdx, a.rdy, a.rdz = g.rdx, g.rdy, g.rdz
        for i, c in enumerate(g.g.face_codes()):
            a.face[i] = c
        if self.material is not None:
            a.has_material = 1
            for i, v in enumerate(self.material):
                a.material[i] = v
        return a

    # field_advance_kernels_t entries on the path (field_advance.h:170-218)
    def advance_b(self, frac):
        _lib.check(_lib.load().vpb_advance_b(C.byref(self.args()), frac, _stream()), "advance_b")

    def advance_e(self, frac=1.0):
        _lib.check(_lib.load().vpb_vacuum_advance_e(C.byref(self.args()), frac, _stream()), "advance_e")

    def clear_jf(self):
        _lib.check(_lib.load().vpb_clear_jf(C.byref(self.args()), _stream()), "clear_jf")

    def synchronize_jf(self):
        _lib.check(_lib.load().vpb_synchronize_jf(C.byref(self.args()), _stream()), "synchronize_jf")

    def energy_f(self):
        _lib.check(_lib.load().vpb_vacuum_energy_f(C.byref(self.args()), _ptr(self._en), _stream()), "energy_f")
        return self._en.cpu().numpy().copy()

    # divergence cleaning and shared-face synchronisation (field_advance.h:186-218, advance.cc:138-176)
    def _call(self, name, *extra):
        _lib.check(getattr(_lib.load(), name)(C.byref(self.args()), *extra, _stream()), name)

    def clear_rhof(self): self._call("vpb_clear_rhof")
    def synchronize_rho(self): self._call("vpb_synchronize_rho")
    def compute_div_e_err(self): self._call("vpb_vacuum_compute_div_e_err")
    def clean_div_e(self): self._call("vpb_vacuum_clean_div_e")
    def compute_div_b_err(self): self._call("vpb_compute_div_b_err")
    def clean_div_b(self): self._call("vpb_clean_div_b")
    def compute_rhob(self): self._call("vpb_vacuum_compute_rhob")
    def compute_curl_b(self): self._call("vpb_vacuum_compute_curl_b")

    def rms_terms(self, name):
        """The two local terms of an rms error, [sum * dV, volume] (compute_rms_div_e_err_pipeline.cc:175-178); the
        caller sums them over ranks and finishes with rms_finish."""
        g = self.g
        self._call(name, _ptr(self._en))
        s = float(self._en[0].item())
        dV = np.float32(g.dV)
        return [s * float(dV), float(np.float32(g.nx * g.ny * g.nz) * dV)]

    def rms_finish(self, terms):
        return float(self.g.eps0) * float(np.sqrt(terms[0] / terms[1]))

    def compute_rms_div_e_err(self): return self.rms_finish(self.rms_terms("vpb_compute_rms_div_e_err"))
    def compute_rms_div_b_err(self): return self.rms_finish(self.rms_terms("vpb_compute_rms_div_b_err"))

    def synchronize_tang_e_norm_b(self, read=True):
        """Local walls and self-periodic axes; the squared-difference sum stays in self._en[0] (read=False) so that a
        slab exchange can add its share before the value is read."""
        self._call("vpb_synchronize_tang_e_norm_b", _ptr(self._en))
        return float(self._en[0].item()) if read else None


class HydroArray:
    """hydro_array_t (sf_interface.h:194-200) on the device: one block, hydro_t = 16 floats per voxel."""

    def __init__(self, g: DeviceGrid):
        self.g = g
        self.h = torch.zeros((g.nv, abi.HYDRO_FLOATS), dtype=torch.float32, device=g.device)

    def clear(self):
        g = self.g
        _lib.check(_lib.load().vpb_clear_hydro(_ptr(self.h), g.nx, g.ny, g.nz, _stream()), "clear_hydro_array")

    def synchronize(self, fa: "FieldArray"):
        """synchronize_hydro_array; the field array only lends its geometry (faces, cell sizes)."""
        _lib.check(_lib.load().vpb_synchronize_hydro(_ptr(self.h), C.byref(fa.args()), _stream()), "synchronize_hydro_array")


def accumulate_hydro_p(ha: HydroArray, sp: "Species", ia: "InterpolatorArray"):
    """accumulate_hydro_p(hydro_array_t*, const species_t*, const interpolator_array_t*), species_advance.h:139-148."""
    _bad_args(ha is None or sp is None or ia is None or ha.g is not sp.g or ha.g is not ia.g, "accumulate_hydro_p")
    g = sp.g
    _lib.check(_lib.load().vpb_accumulate_hydro_p(_ptr(ha.h), _ptr(sp.p), sp.np, _ptr(ia.i), ia.stride,
                                                  sp.q, sp.m, g.dt, g.cvac, g.g.r8V, g.nx, g.ny, g.nz, _stream()),
               "accumulate_hydro_p")


class Species:
    """species_t (species_advance_aos.h:54-94) with device arrays."""

    def __init__(self, name, q, m, max_np, max_nm, sort_interval, sort_out_of_place, g: DeviceGrid):
        _bad_args(not name or g is None, "species")
        self.name, self.q, self.m = name, float(np.float32(q)), float(np.float32(m))
        self.max_np, self.max_nm = max(1, int(max_np)), max(1, int(max_nm))
        self.np, self.nm = 0, 0
        self.g = g
        self.sort_interval, self.sort_out_of_place = sort_interval, sort_out_of_place
        self.last_sorted = -(2 ** 63)
        dev = g.device
        self._p = torch.zeros((self.max_np, 8), dtype=torch.float32, device=dev)
        self._perm = None                 # order of a deferred sort_p the next advance_p applies (sort_p(sp, defer=True))
        self._perm_pending = False
        self._keys = None                 # voxel of every particle as the last advance_p(emit_keys=True) left it ...
        self._keys_valid = False          # ... still true: nothing has touched the array since
        self._keys_np = -1
        self.pm = torch.zeros((self.max_nm, 4), dtype=torch.float32, device=dev)
        self.partition = torch.zeros(g.nv + 1, dtype=torch.int32, device=dev)
        self.counters = torch.zeros(4, dtype=torch.int32, device=dev)
        self.n_ignored = 0
        self.partition_np = -1            # particle count of the last sort_p (partition[] valid), -1 = never sorted
        self._aux = None
        self._scratch = None
        self._mv_scratch = None
        self._en = torch.zeros(1, dtype=torch.float64, device=dev)

    @property
    def p(self):
        """particle_t array [max_np, 8 floats].  Reading it settles a deferred sort first, so every user except the
        advance_p that consumes the order sees the array sort_p promised."""
        if self._perm_pending:
            self.settle()
        self._keys_valid = False          # whoever holds the tensor may write it
        return self._p

    def settle(self):
        """Apply a pending sort_p order now (vpb_permute_p into the aux array, then swap)."""
        if not self._perm_pending:
            return
        self._perm_pending = False
        self._keys_valid = False
        _lib.check(_lib.load().vpb_permute_p(_ptr(self._p), self.np, _ptr(self._perm), _ptr(self._aux), _stream()), "permute_p")
        self._p, self._aux = self._aux, self._p

    def set_particles(self, arr):
        """arr: numpy structured array (abi.particle_dtype) or float32 [n,8] tensor."""
        if isinstance(arr, np.ndarray):
            t = torch.from_numpy(arr.view(np.float32).reshape(-1, 8))
        else:
            t = arr
        n = t.shape[0]
        _bad_args(n > self.max_np, "set_particles")
        self._perm_pending = False
        self._keys_valid = False
        self._p[:n].copy_(t)
        self.np, self.nm = n, 0
        self.partition_np = -1

    def particles_host(self):
        return self.p[:self.np].cpu().numpy().reshape(-1).view(abi.particle_dtype)

    def movers_host(self):
        return self.pm[:self.nm].cpu().numpy().reshape(-1).view(abi.mover_dtype)

    def push_constants(self):
        """advance_p_pipeline.cc:279-283, evaluated in float like the reference does."""
        f32, g = np.float32, self.g
        q, m, dt, cvac = f32(self.q), f32(self.m), f32(g.dt), f32(g.cvac)
        return (f32(f32(q * dt) / f32(f32(f32(2) * m) * cvac)),
                f32(f32(cvac * dt) * f32(g.rdx)), f32(f32(cvac * dt) * f32(g.rdy)), f32(f32(cvac * dt) * f32(g.rdz)), q)


# ---- operators (same names as the reference's C API) ---------------------------------------------------------

def advance_p(sp: Species, aa: AccumulatorArray, ia: InterpolatorArray, variant=_lib.DEPOSIT_DEFAULT, sync=True,
              emit_keys=False):
    """advance_p(species_t*, accumulator_array_t*, const interpolator_array_t*), species_advance.h:73-76.
    emit_keys=True (the caller knows a sort_p comes next): the push also leaves every particle's voxel in a compact
    array, so that sort does not have to read the particles to find it."""
    _bad_args(sp is None or aa is None or ia is None or sp.g is not aa.g or sp.g is not ia.g, "advance_p")
    g = sp.g
    sp.counters.zero_()
    a = _lib.PushArgs()
    gather = sp._perm_pending and variant in (_lib.DEPOSIT_DEFAULT, 2) and not int(os.environ.get("VPB_DEBUG_SKIP", "0"))
    if gather:                            # the push applies the order of the deferred sort_p: p[perm[k]] -> aux[k]
        a.p, a.np = sp._p.data_ptr(), sp.np
        a.perm, a.p_out = sp._perm.data_ptr(), sp._aux.data_ptr()
    else:
        sp.settle()
        a.p, a.np = sp._p.data_ptr(), sp.np
    sp._keys_valid = False
    emit_keys = emit_keys and variant in (_lib.DEPOSIT_DEFAULT, 2) and not int(os.environ.get("VPB_DEBUG_SKIP", "0"))
    if emit_keys:
        if sp._keys is None:
            sp._keys = torch.empty(sp.max_np, dtype=torch.int32, device=sp.g.device)
        a.keys_out = sp._keys.data_ptr()
    a.pm, a.max_nm = sp.pm.data_ptr(), sp.max_nm
    a.counters = sp.counters.data_ptr()
    a.interp, a.interp_stride = ia.i.data_ptr(), ia.stride
    a.accum, a.accum_stride = aa.a.data_ptr(), aa.stride_floats
    a.neighbor, a.rangel, a.rangeh = g.neighbor.data_ptr(), g.rangel, g.rangeh
    a.qdt_2mc, a.cdt_dx, a.cdt_dy, a.cdt_dz, a.qsp = sp.push_constants()
    a.nx, a.ny, a.nz = g.nx, g.ny, g.nz
    a.variant = variant
    rule = g.neighbor_rule()
    if rule is not None:
        a.neighbor_rule = rule
    a.debug_skip = int(os.environ.get("VPB_DEBUG_SKIP", "0"))
    if sp.partition_np >= 0:              # partition[] of the last sort_p: advance_p walks the array brick by brick
        a.partition, a.partition_np = sp.partition.data_ptr(), sp.partition_np
    L = _lib.load()
    _lib.check(L.vpb_advance_p(C.byref(a), _stream()), "advance_p")
    if gather:
        sp._perm_pending = False
        sp._p, sp._aux = sp._aux, sp._p
    sp._keys_valid = bool(emit_keys)      # void again if any mover left the domain (checked by sort_p through sp.nm)
    sp._keys_np = sp.np
    if sync:
        finish_advance_p(sp)


def finish_advance_p(sp: Species):
    """Read back the mover count (sp->nm) and put the movers in ascending particle order for boundary_p."""
    finish_advance_p_all([sp])


def finish_advance_p_all(species):
    """finish_advance_p for several species with a single device->host read of all their counters."""
    if not species:
        return
    L = _lib.load()
    c = torch.stack([sp.counters for sp in species]).cpu()
    for k, sp in enumerate(species):
        sp.nm = min(int(c[k, 0]), sp.max_nm)
        sp.n_ignored = int(c[k, 1])
        if sp.n_ignored:
            import warnings
            warnings.warn(f"species {sp.name} ran out of storage for {sp.n_ignored} movers")
        if sp.nm > 1:
            need = L.vpb_sort_movers_scratch_bytes(sp.nm)
            if sp._mv_scratch is None or sp._mv_scratch.numel() < need:
                sp._mv_scratch = torch.empty(int(need * 1.5), dtype=torch.uint8, device=sp.g.device)
            _lib.check(L.vpb_sort_movers(_ptr(sp.pm), sp.nm, _ptr(sp._mv_scratch), sp._mv_scratch.numel(), _stream()),
                       "sort_movers")


def sort_p(sp: Species, defer=False):
    """sort_p(species_t*), species_advance.h:65-66.  defer=True computes the same order and partition[] from 8-byte
    (voxel, index) pairs and leaves the particles where they are; the next advance_p(sp, ...) moves them while it pushes
    them (one pass over the particle array instead of five), anything else that looks at sp.p settles the order first."""
    _bad_args(sp is None, "sort_p")
    L = _lib.load()
    g = sp.g
    sp.settle()
    sp.last_sorted = g.step
    if sp._aux is None or sp._aux.shape[0] < sp.np:
        sp._aux = torch.empty((sp.max_np, 8), dtype=torch.float32, device=g.device)
    if defer and sp.np > 1:
        need = L.vpb_sort_index_scratch_bytes(max(sp.np, 1), g.nv)
        if sp._scratch is None or sp._scratch.numel() < need:
            sp._scratch = torch.empty(need, dtype=torch.uint8, device=g.device)
        if sp._perm is None:
            sp._perm = torch.empty(sp.max_np, dtype=torch.int32, device=g.device)
        work_bytes = sp._aux.numel() * 4
        _bad_args(work_bytes < L.vpb_sort_index_work_bytes(sp.np), "sort_p")
        # the voxels the last push left behind, if nothing has touched the array since (no boundary_p, no host access)
        keys = _ptr(sp._keys) if (sp._keys_valid and sp.nm == 0 and sp._keys_np == sp.np) else None
        sp._keys_valid = False
        _lib.check(L.vpb_sort_p_index(_ptr(sp._p), keys, sp.np, _ptr(sp._perm), _ptr(sp.partition), g.nx, g.ny, g.nz,
                                      _ptr(sp._aux), work_bytes, _ptr(sp._scratch), sp._scratch.numel(), _stream(), None),
                   "sort_p")
        sp._perm_pending = True
    else:
        need = L.vpb_sort_scratch_bytes(max(sp.np, 1), g.nv)
        if sp._scratch is None or sp._scratch.numel() < need:
            sp._scratch = torch.empty(need, dtype=torch.uint8, device=g.device)
        _lib.check(L.vpb_sort_p(_ptr(sp._p), sp.np, _ptr(sp._aux), _ptr(sp.partition), g.nx, g.ny, g.nz,
                                _ptr(sp._scratch), sp._scratch.numel(), _stream()), "sort_p")
    sp.partition_np = sp.np


def load_interpolator_array(ia: InterpolatorArray, fa: FieldArray):
    _bad_args(ia is None or fa is None or ia.g is not fa.g, "load_interpolator_array")
    g = ia.g
    _lib.check(_lib.load().vpb_load_interpolator(_ptr(ia.i), ia.stride, _ptr(fa.f), g.nx, g.ny, g.nz, _stream()),
               "load_interpolator_array")


def clear_accumulator_array(aa: AccumulatorArray):
    _bad_args(aa is None, "clear_accumulator_array")
    g = aa.g
    _lib.check(_lib.load().vpb_clear_accumulator(_ptr(aa.a), aa.stride_floats, g.nx, g.ny, g.nz, _stream()),
               "clear_accumulator_array")


def reduce_accumulator_array(aa: AccumulatorArray):
    """Identity on the device: there is a single accumulator block (the reference sums n_pipeline+1 blocks)."""
    _bad_args(aa is None, "reduce_accumulator_array")


def unload_accumulator_array(fa: FieldArray, aa: AccumulatorArray):
    _bad_args(fa is None or aa is None or fa.g is not aa.g, "unload_accumulator_array")
    g = fa.g
    _lib.check(_lib.load().vpb_unload_accumulator(_ptr(fa.f), _ptr(aa.a), aa.stride_floats, g.nx, g.ny, g.nz,
                                                  g.rdx, g.rdy, g.rdz, g.dt, _stream()), "unload_accumulator_array")


def energy_p(sp: Species, ia: InterpolatorArray):
    _bad_args(sp is None or ia is None or sp.g is not ia.g, "energy_p")
    g = sp.g
    _lib.check(_lib.load().vpb_energy_p(_ptr(sp.p), sp.np, _ptr(ia.i), ia.stride, sp.q, sp.m, g.dt, g.cvac,
                                        _ptr(sp._en), _stream()), "energy_p")
    return float(sp._en.cpu()[0])


def accumulate_rho_p(fa: FieldArray, sp: Species):
    """accumulate_rho_p(field_array_t*, const species_t*), species_advance.h:117-119."""
    _bad_args(fa is None or sp is None or fa.g is not sp.g, "accumulate_rho_p")
    g = sp.g
    _lib.check(_lib.load().vpb_accumulate_rho_p(_ptr(fa.f), _ptr(sp.p), sp.np, sp.q, g.g.r8V, g.nx, g.ny, g.nz, _stream()),
               "accumulate_rho_p")


def center_p(sp: Species, ia: InterpolatorArray):
    _bad_args(sp is None or ia is None or sp.g is not ia.g, "center_p")
    _lib.check(_lib.load().vpb_center_p(_ptr(sp.p), sp.np, _ptr(ia.i), ia.stride, sp.push_constants()[0], _stream()),
               "center_p")


def uncenter_p(sp: Species, ia: InterpolatorArray):
    _bad_args(sp is None or ia is None or sp.g is not ia.g, "uncenter_p")
    _lib.check(_lib.load().vpb_uncenter_p(_ptr(sp.p), sp.np, _ptr(ia.i), ia.stride, sp.push_constants()[0], _stream()),
               "uncenter_p")


# ---- boundary_p, particle side (src/boundary/boundary_p.cc:257-371,595-711) ----------------------------------

def boundary_pack(sp: Species, face_range, fa: FieldArray = None, absorb_all=False):
    """Turn this species' movers into per-face injector buffers and back-fill the holes they leave.

    Returns (inj, offsets_dev): inj is a [nm, 12] float32 view of particle_injector_t records grouped by class,
    offsets_dev an int32[9] device tensor (class c occupies [offsets[c], offsets[c+1]); 0..5 faces, 6 absorbed,
    7 no handler).  sp.np and sp.nm are updated (every mover's particle leaves the array)."""
    L = _lib.load()
    g = sp.g
    offs = torch.zeros(9, dtype=torch.int32, device=g.device)
    nm = sp.nm
    if nm == 0:
        return None, offs
    inj = torch.empty((nm, 12), dtype=torch.float32, device=g.device)
    need = L.vpb_boundary_scratch_bytes(nm)
    scratch = torch.empty(need, dtype=torch.uint8, device=g.device)
    b = _lib.BoundaryArgs()
    b.p, b.np, b.pm, b.nm = sp.p.data_ptr(), sp.np, sp.pm.data_ptr(), nm
    b.neighbor = g.neighbor.data_ptr()
    b.rangel, b.rangeh, b.rangem = g.rangel, g.rangeh, int(g.g.range[g.world_size])
    for f in range(6):
        b.face_range[f] = face_range[f]
    b.sp_id = getattr(sp, "id", 0)
    b.inj, b.class_offsets = inj.data_ptr(), offs.data_ptr()
    b.scratch, b.scratch_bytes = scratch.data_ptr(), need
    b.absorb_all = 1 if absorb_all else 0
    if fa is not None:                                   # absorbed particles leave their charge in rhob
        b.fields, b.q_r8V = fa.f.data_ptr(), float(np.float32(np.float32(sp.q) * np.float32(g.g.r8V)))
        b.nx, b.ny, b.nz = g.nx, g.ny, g.nz
    _lib.check(L.vpb_boundary_p_pack(C.byref(b), _stream()), "boundary_p_pack")
    sp.np -= nm
    sp.nm = 0
    sp._bp_keep = (inj, scratch)          # keep alive until the stream has consumed them
    return inj, offs


def boundary_inject(sp: Species, aa: AccumulatorArray, ia: InterpolatorArray, inj, n):
    """Append n received injectors (device tensor [n,12]) and finish their moves; new movers accumulate in sp.pm
    through sp.counters (call finish_advance_p afterwards to read sp.nm)."""
    if n == 0:
        return
    if sp.np + n > sp.max_np:
        raise RuntimeError(f"species {sp.name}: {sp.np}+{n} particles exceed max_np={sp.max_np}")
    g = sp.g
    a = _lib.PushArgs()
    a.p, a.np = sp.p.data_ptr(), sp.np
    a.pm, a.max_nm = sp.pm.data_ptr(), sp.max_nm
    a.counters = sp.counters.data_ptr()
    a.interp, a.interp_stride = ia.i.data_ptr(), ia.stride
    a.accum, a.accum_stride = aa.a.data_ptr(), aa.stride_floats
    a.neighbor, a.rangel, a.rangeh = g.neighbor.data_ptr(), g.rangel, g.rangeh
    a.qdt_2mc, a.cdt_dx, a.cdt_dy, a.cdt_dz, a.qsp = sp.push_constants()
    a.nx, a.ny, a.nz = g.nx, g.ny, g.nz
    rule = g.neighbor_rule()
    if rule is not None:
        a.neighbor_rule = rule
    _lib.check(_lib.load().vpb_boundary_p_inject(C.byref(a), _ptr(inj), n, _stream()), "boundary_p_inject")
    sp.np += n


# ---- fixed-capacity migration messages (count in the header; no host synchronisation inside a round) ----------

def boundary_msg_floats(cap):
    """float32 elements of a migration message of capacity `cap` (16-byte header + cap particle_injector_t)."""
    return 4 + 12 * int(cap)


def boundary_stage(inj, offs, face, cap, sp_id, msg, status):
    """Copy class `face` of a boundary_pack result into the fixed-capacity message `msg` and write its header."""
    _lib.check(_lib.load().vpb_boundary_p_stage(_ptr(inj), _ptr(offs), face, cap, sp_id, _ptr(msg), _ptr(status), _stream()),
               "boundary_p_stage")


def boundary_inject_msg(sp: Species, aa: AccumulatorArray, ia: InterpolatorArray, msg, cap, added, status):
    """Append the records of a received message behind p[sp.np + added) (device-side count) and finish their moves."""
    g = sp.g
    a = _lib.PushArgs()
    a.p, a.np = sp.p.data_ptr(), sp.np
    a.pm, a.max_nm = sp.pm.data_ptr(), sp.max_nm
    a.counters = sp.counters.data_ptr()
    a.interp, a.interp_stride = ia.i.data_ptr(), ia.stride
    a.accum, a.accum_stride = aa.a.data_ptr(), aa.stride_floats
    a.neighbor, a.rangel, a.rangeh = g.neighbor.data_ptr(), g.rangel, g.rangeh
    a.qdt_2mc, a.cdt_dx, a.cdt_dy, a.cdt_dz, a.qsp = sp.push_constants()
    a.nx, a.ny, a.nz = g.nx, g.ny, g.nz
    rule = g.neighbor_rule()
    if rule is not None:
        a.neighbor_rule = rule
    _lib.check(_lib.load().vpb_boundary_p_inject_msg(C.byref(a), _ptr(msg), cap, sp.max_np, _ptr(added), _ptr(status), _stream()),
               "boundary_p_inject_msg")


def drop_unresolved_movers(sp: Species, fa: FieldArray):
    """Movers still unresolved after the last communication round (src/vpic/advance.cc:78-101): the reference warns,
    puts their charge into rhob and removes them from the particle array; so does this."""
    if sp.nm == 0:
        return 0
    import warnings
    n = sp.nm
    warnings.warn(f"Ignoring {n} unprocessed {sp.name} movers (increase num_comm_round)")
    boundary_pack(sp, [-1] * 6, fa, absorb_all=True)
    return n

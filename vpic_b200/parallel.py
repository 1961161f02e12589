"""Multi-GPU plumbing for slab-decomposed runs: one process per GPU, torch.distributed (NCCL over NVLink) for the
neighbour exchanges that the reference does with MPI Issend/Irecv over its 6-face ports (src/util/mp/DMPPolicy.h,
src/grid/grid_comm.cc):

  boundary_p      particle migration, src/boundary/boundary_p.cc:392-446 (counts, then particle_injector_t payloads)
  synchronize_jf  shared-plane current sums, src/field_advance/standard/remote.cc:417-508
  ghost_tang_b    tangential-B ghost planes inside advance_e, remote.cc:61-134

A slab decomposition along one periodic axis gives every GPU exactly two neighbours (possibly the same rank twice
when world_size == 2).  All collectives here are neighbour send/recv — the path has no all-to-all.  Packing,
unpacking, back-fill and injection are CUDA kernels (boundary_p.cu, field_advance.cu); this module only routes
buffers.  The same routing runs over gloo with CPU tensors in the CPU test-suite.
"""
import os

import torch
import torch.distributed as dist

from . import engine as E, lib as _lib


class NeighbourRing:
    """Send one buffer to each neighbour along the slab axis and receive theirs.

    Message matching: a batch posts [send->lo, send->hi, recv<-hi, recv<-lo].  With distinct neighbours order is
    irrelevant; when both neighbours are the same rank (world_size == 2) the peer's first receive (from its `hi`,
    i.e. us) is matched by our first send (to `lo`, i.e. the peer) — what leaves through our low face must arrive
    through the peer's high face."""

    def __init__(self, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi = (rank - 1) % world, (rank + 1) % world

    def sendrecv(self, out_lo, out_hi, in_lo, in_hi):
        """out_lo goes to the low neighbour (it receives it as its in_hi), out_hi to the high neighbour."""
        if self.world == 1:
            in_hi.copy_(out_lo)
            in_lo.copy_(out_hi)
            return
        ops = []
        if out_lo.numel():
            ops.append(dist.P2POp(dist.isend, out_lo, self.lo, self.group))
        if out_hi.numel():
            ops.append(dist.P2POp(dist.isend, out_hi, self.hi, self.group))
        if in_hi.numel():
            ops.append(dist.P2POp(dist.irecv, in_hi, self.hi, self.group))
        if in_lo.numel():
            ops.append(dist.P2POp(dist.irecv, in_lo, self.lo, self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()


class SlabExchange:
    def __init__(self, dgrid: E.DeviceGrid, axis=1, group=None):
        self.g = dgrid
        self.axis = axis
        self.f_lo, self.f_hi = axis, axis + 3
        self.ring = NeighbourRing(dgrid.rank, dgrid.world_size, group)
        rng = dgrid.g.range
        self.face_range = [-1] * 6
        self.face_range[self.f_lo] = int(rng[self.ring.lo])
        self.face_range[self.f_hi] = int(rng[self.ring.hi])
        n = _lib.load().vpb_halo_floats(dgrid.nx, dgrid.ny, dgrid.nz, axis)
        dev = dgrid.device
        self.h_out = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
        self.h_in = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
        self._extra = {}
        # fixed-capacity migration (boundary_p_fixed): per-species message buffers and deferred results
        self.fixed = os.environ.get("VPB_EXCHANGE_FIXED", "1") != "0"
        self.calibration_steps = 2          # steps that use the counted exchange and measure the message sizes
        self._steps_seen = 0
        self._mig = {}
        self._pending = None
        # After an injection a particle can become a mover again only at a wall that absorbs (or a custom handler):
        # every other face of a slab is periodic-self, reflecting or one of the two shared ones it just came through.
        nb = getattr(dgrid.g, "neighbor", None)                  # particle boundary codes live in grid_t.neighbor (< 0)
        self.inject_may_emit = bool((nb <= -2).any()) if nb is not None else False

    # ---- deferred bookkeeping of the fixed-capacity exchange ------------------------------------------------------
    def begin_step(self, sim):
        self.resolve(sim)

    def resolve(self, sim):
        """Read what the last fixed-capacity round left on the device: particles appended per species, capacity and
        array overflows, and the largest message anywhere in the ring (which sizes the next step's messages)."""
        if self._pending is None:
            return
        ev, host, sps = self._pending
        self._pending = None
        ev.synchronize()
        vals = host.tolist()
        gmax = vals[-1]
        for k, sp in enumerate(sps):
            added, status, emitted = vals[3 * k], vals[3 * k + 1], vals[3 * k + 2]
            sp.np += added
            if status & 1:
                raise RuntimeError(f"species {sp.name}: a migration message exceeded its capacity of {self._mig[sp.id]['cap']} "
                                   "particles (set VPB_EXCHANGE_FIXED=0 for the counted exchange)")
            if status & 2:
                raise RuntimeError(f"species {sp.name}: injected particles exceed max_np={sp.max_np}")
            if status & 4:
                raise RuntimeError(f"species {sp.name}: particles hit a boundary with no device handler (custom particle "
                                   "boundary conditions stay on the host)")
            if emitted and not self.inject_may_emit:
                raise RuntimeError(f"species {sp.name}: {emitted} injected particles became movers again although no wall "
                                   "of this slab can emit them")
        # every rank sees the same ring-wide maximum, so every rank picks the same capacity for the next step
        for sp in sps:
            m = self._mig[sp.id]
            if gmax > m["cap"] // 2:
                self._alloc(sp, 4 * gmax)

    def _alloc(self, sp, cap):
        cap = max(4096, (int(cap) + 1023) // 1024 * 1024)
        dev = self.g.device
        n = E.boundary_msg_floats(cap)
        old = self._mig.get(sp.id)
        self._mig[sp.id] = {
            "cap": cap,
            "out": [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(2)],
            "in": [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(2)],
            "added": old["added"] if old else torch.zeros(1, dtype=torch.int32, device=dev),
            "status": old["status"] if old else torch.zeros(2, dtype=torch.int32, device=dev),
        }

    def boundary_p_fixed(self, sim, species):
        """First communication round of boundary_p for `species` with fixed-capacity messages: pack, stage, one batch
        of neighbour sends and receives whose sizes both sides know in advance, inject.  No device-to-host read and no
        collective: the counts ride in the message headers (boundary_p.cc:205-211 reserves the same header) and are
        consumed by the injection kernel; sp.np catches up at the next resolve()."""
        for sp in species:
            m = self._mig[sp.id]
            cap = m["cap"]
            inj, offs = E.boundary_pack(sp, self.face_range, sim.field_array)
            m["status"][0:1].bitwise_or_(((offs[8] - offs[7]) > 0).to(torch.int32).reshape(1) * 4)   # class 7: no device handler
            E.boundary_stage(inj, offs, self.f_lo, cap, sp.id, m["out"][0], m["status"])
            E.boundary_stage(inj, offs, self.f_hi, cap, sp.id, m["out"][1], m["status"])
            self.ring.sendrecv(m["out"][0], m["out"][1], m["in"][0], m["in"][1])
            sp.counters.zero_()
            m["added"].zero_()
            # injection order of the reference: faces 0..5, so the low face first
            E.boundary_inject_msg(sp, sim.accumulator_array, sim.interpolator_array, m["in"][0], cap, m["added"], m["status"])
            E.boundary_inject_msg(sp, sim.accumulator_array, sim.interpolator_array, m["in"][1], cap, m["added"], m["status"])

    def end_fixed(self, sim, species):
        """Queue the read-back of the step's device-side results (one small copy, resolved at the next step)."""
        dev = self.g.device
        parts = []
        for sp in species:
            m = self._mig[sp.id]
            parts += [m["added"], m["status"][0:1], sp.counters[0:1]]
        sizes = torch.stack([self._mig[sp.id]["status"][1] for sp in species]).max().reshape(1)
        if self.ring.world > 1:
            dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=self.ring.group)
        flat = torch.cat(parts + [sizes])
        host = torch.empty(flat.shape, dtype=torch.int32, pin_memory=(dev.type == "cuda"))
        host.copy_(flat, non_blocking=True)
        for sp in species:
            self._mig[sp.id]["status"].zero_()
        ev = torch.cuda.Event() if dev.type == "cuda" else _NoEvent()
        ev.record()
        self._pending = (ev, host, list(species))

    def use_fixed(self, sim):
        """Counted exchange for the first steps (they measure the message sizes), fixed-capacity messages afterwards."""
        if not self.fixed:
            return False
        if self._steps_seen < self.calibration_steps:
            return False
        if not self._mig:
            cmax = torch.tensor([self._calib_max], dtype=torch.int32, device=self.g.device)
            if self.ring.world > 1:
                dist.all_reduce(cmax, op=dist.ReduceOp.MAX, group=self.ring.group)
            for sp in sim.species_list:
                self._alloc(sp, 4 * int(cmax.item()))
        return True

    _calib_max = 0

    # ---- particles ------------------------------------------------------------------------------------------
    def boundary_p(self, sim, species=None, check_empty=True):
        """One communication round of boundary_p for `species` (default: every species; the caller loops
        num_comm_round times).

        A round in which no rank holds a mover is a no-op in the reference too (zero counts both ways); one tiny
        all-reduce finds that out, so the usual second and third rounds cost one collective instead of a full
        count/payload handshake."""
        dev = self.g.device
        sps = sim.species_list if species is None else species
        if self.ring.world > 1 and check_empty:
            left = torch.tensor([sum(sp.nm for sp in sps)], dtype=torch.int32, device=dev)
            dist.all_reduce(left, group=self.ring.group)
            if int(left.item()) == 0:
                return False                      # nobody holds a mover: this round and every later one is a no-op
        packed = [E.boundary_pack(sp, self.face_range, sim.field_array) for sp in sps]
        offs = torch.stack([o for _, o in packed]).cpu()                       # one sync for all species
        n_lo = [int(offs[s, self.f_lo + 1] - offs[s, self.f_lo]) for s in range(len(sps))]
        n_hi = [int(offs[s, self.f_hi + 1] - offs[s, self.f_hi]) for s in range(len(sps))]
        self._calib_max = max([self._calib_max] + n_lo + n_hi)
        for s, sp in enumerate(sps):
            if int(offs[s, 8] - offs[s, 7]):
                raise RuntimeError(f"species {sp.name}: {int(offs[s, 8] - offs[s, 7])} particles hit a boundary with "
                                   "no device handler (custom particle boundary conditions stay on the host)")
        c_out_lo = torch.tensor(n_lo, dtype=torch.int32, device=dev)
        c_out_hi = torch.tensor(n_hi, dtype=torch.int32, device=dev)
        c_in_lo, c_in_hi = torch.empty_like(c_out_lo), torch.empty_like(c_out_hi)
        self.ring.sendrecv(c_out_lo, c_out_hi, c_in_lo, c_in_hi)
        r_lo, r_hi = c_in_lo.cpu().tolist(), c_in_hi.cpu().tolist()
        empty = torch.empty((0, 12), dtype=torch.float32, device=dev)
        for s, sp in enumerate(sps):
            inj = packed[s][0]
            o = offs[s]
            out_lo = inj[int(o[self.f_lo]):int(o[self.f_lo + 1])] if n_lo[s] else empty
            out_hi = inj[int(o[self.f_hi]):int(o[self.f_hi + 1])] if n_hi[s] else empty
            in_lo = torch.empty((r_lo[s], 12), dtype=torch.float32, device=dev)
            in_hi = torch.empty((r_hi[s], 12), dtype=torch.float32, device=dev)
            # the peer sizes its receive from the count we sent, so empty messages are skipped on both sides
            self.ring.sendrecv(out_lo, out_hi, in_lo, in_hi)
            sp.counters.zero_()
            # injection order of the reference: faces 0..5, so the low face first
            E.boundary_inject(sp, sim.accumulator_array, sim.interpolator_array, in_lo, r_lo[s])
            E.boundary_inject(sp, sim.accumulator_array, sim.interpolator_array, in_hi, r_hi[s])
        E.finish_advance_p_all(sps)
        return True

    # ---- fields ---------------------------------------------------------------------------------------------
    def _buffers(self, kind):
        """Send/receive planes for one halo kind (TANG_B and JF share the preallocated pair)."""
        if kind in (_lib.HALO_TANG_B, _lib.HALO_JF):
            return self.h_out, self.h_in
        if kind not in self._extra:
            g = self.g
            n = _lib.load().vpb_halo_floats_kind(g.nx, g.ny, g.nz, self.axis, kind)
            mk = lambda: [torch.empty(n, dtype=torch.float32, device=g.device) for _ in range(2)]
            self._extra[kind] = (mk(), mk())
        return self._extra[kind]

    def _halo(self, fa, kind, err=None):
        import ctypes as C
        L = _lib.load()
        a = fa.args()
        st = E._stream()
        out, inn = self._buffers(kind)
        _lib.check(L.vpb_halo_pack(C.byref(a), kind, self.f_lo, E._ptr(out[0]), st), "halo_pack")
        _lib.check(L.vpb_halo_pack(C.byref(a), kind, self.f_hi, E._ptr(out[1]), st), "halo_pack")
        self.ring.sendrecv(out[0], out[1], inn[0], inn[1])
        if kind == _lib.HALO_TANG_E_NORM_B:
            _lib.check(L.vpb_halo_unpack_sync(C.byref(a), self.f_lo, E._ptr(inn[0]), E._ptr(err), st), "halo_unpack_sync")
            _lib.check(L.vpb_halo_unpack_sync(C.byref(a), self.f_hi, E._ptr(inn[1]), E._ptr(err), st), "halo_unpack_sync")
            return
        _lib.check(L.vpb_halo_unpack(C.byref(a), kind, self.f_lo, E._ptr(inn[0]), st), "halo_unpack")
        _lib.check(L.vpb_halo_unpack(C.byref(a), kind, self.f_hi, E._ptr(inn[1]), st), "halo_unpack")

    def synchronize_jf(self, sim):
        self._halo(sim.field_array, _lib.HALO_JF)

    def ghost_tang_b(self, sim):
        self._halo(sim.field_array, _lib.HALO_TANG_B)

    # the periodic extras of divergence cleaning and shared-face synchronisation (remote.cc:136-416,534-620)
    def synchronize_rho(self, sim):
        self._halo(sim.field_array, _lib.HALO_RHO)

    def ghost_norm_e(self, sim):
        self._halo(sim.field_array, _lib.HALO_NORM_E)

    def ghost_div_b(self, sim):
        self._halo(sim.field_array, _lib.HALO_DIV_B)

    def synchronize_tang_e_norm_b(self, sim, err_dev):
        """Average the shared planes with the neighbours'; the squared differences are added to err_dev[0]."""
        self._halo(sim.field_array, _lib.HALO_TANG_E_NORM_B, err=err_dev)

    def allsum(self, values):
        """mp_allsum_d (src/util/mp/mp.h) for a few host doubles."""
        t = torch.tensor(values, dtype=torch.float64, device=self.g.device)
        if self.ring.world > 1:
            dist.all_reduce(t, group=self.ring.group)
        return t.cpu().tolist()


class _NoEvent:
    """CPU stand-in for torch.cuda.Event (the gloo tests run the same bookkeeping without a device)."""
    def record(self): pass
    def synchronize(self): pass

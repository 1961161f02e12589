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

    def begin_step(self, sim):
        pass

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

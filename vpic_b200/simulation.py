"""The hot-path block of vpic_simulation::advance() (src/vpic/advance.cc:25-185) over device-resident arrays.

Order of operations is the reference's:
  sort_p (per species, every sort_interval)            advance.cc:25-29
  clear_accumulator_array                              :36-37
  advance_p (per species)                              :49-50
  reduce_accumulator_array                             :65-66
  boundary_p x num_comm_round                          :73-77   (slab-decomposed runs: NCCL, see parallel.py)
  clear_jf ; unload_accumulator_array ; synchronize_jf :107-110
  advance_b(1/2) ; advance_e(1) ; advance_b(1/2)       :123-137
  clean_div_e / clean_div_b / synchronize_tang_e_norm_b at their intervals   :138-176
  load_interpolator_array                              :185
Host hooks of the reference (user_* injections, collisions, emitters, dumps) are outside the hot path and are not
called here.
"""
import os

import torch

from . import engine as E


class Simulation:
    def __init__(self, dgrid: E.DeviceGrid, damp=0.0, simd_width=4, num_comm_round=3, exchange=None):
        self.g = dgrid
        self.field_array = E.FieldArray(dgrid, damp=damp)
        self.interpolator_array = E.InterpolatorArray(dgrid, simd_width)
        self.accumulator_array = E.AccumulatorArray(dgrid, simd_width)
        self.species_list = []
        self.num_comm_round = num_comm_round
        self.exchange = exchange          # parallel.SlabExchange for multi-GPU runs, None on one GPU
        self.deposit_variant = 0
        # sort_p computes the order only and the advance_p that follows moves the particles while it pushes them
        # (engine.sort_p(defer=True)); VPB_DEFER_SORT=0 restores the stand-alone sort
        self.defer_sort = os.environ.get("VPB_DEFER_SORT", "1") != "0"
        self.push_events = None           # list of (start, end) CUDA events around each advance_p when profiling
        self.overlap_exchange = True      # multi-GPU: migrate species k while species k+1 is pushed
        self._side = None
        # divergence cleaning and shared-face synchronisation (vpic.h: clean_div_e_interval, num_div_e_round, ...);
        # 0 = off, as in a deck that does not set them
        self.clean_div_e_interval = 0
        self.clean_div_b_interval = 0
        self.sync_shared_interval = 0
        self.num_div_e_round = 2
        self.num_div_b_round = 2
        self.cleaning_log = []            # (step, what, value): the rms errors / desynchronisation the reference prints

    def define_species(self, name, q, m, max_np, max_nm, sort_interval=20, sort_out_of_place=0):
        sp = E.Species(name, q, m, max_np, max_nm, sort_interval, sort_out_of_place, self.g)
        sp.id = len(self.species_list)
        self.species_list.append(sp)
        return sp

    @property
    def step(self):
        return self.g.g.step

    def initialize(self):
        """The part of vpic_simulation::initialize on the path (initialize.cc:52): interpolators from the fields."""
        if self.exchange is not None:
            self.exchange.begin_step(self)
        E.load_interpolator_array(self.interpolator_array, self.field_array)

    def sync_counts(self):
        """Bring sp.np up to date with what the last migration round appended on the device (multi-GPU runs defer that
        read to the next step)."""
        if self.exchange is not None:
            self.exchange.resolve(self)

    def advance(self):
        fa, ia, aa = self.field_array, self.interpolator_array, self.accumulator_array
        step = self.step
        self.sync_counts()
        for sp in self.species_list:
            if sp.sort_interval > 0 and step % sp.sort_interval == 0:
                E.sort_p(sp, defer=self.defer_sort)
        E.clear_accumulator_array(aa)
        done = []
        for sp in self.species_list:
            if self.push_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            # a sort_p of this species opens the next step: let the push leave the voxel keys behind for it
            keys = self.defer_sort and sp.sort_interval > 0 and (step + 1) % sp.sort_interval == 0
            E.advance_p(sp, aa, ia, variant=self.deposit_variant, sync=False, emit_keys=keys)
            if self.push_events is not None:
                e1.record()
                self.push_events.append((e0, e1, sp.np))
            if self.exchange is not None:
                ev = torch.cuda.Event()
                ev.record()
                done.append(ev)
        if self.exchange is not None and self.overlap_exchange:
            # First migration round per species on a side stream: species k's movers are packed, exchanged over NCCL
            # and injected while species k+1 is still being pushed (injection deposits are atomics, like the push's).
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(priority=-1)
            fixed = self.exchange.use_fixed(self)
            with torch.cuda.stream(self._side):
                for sp, ev in zip(self.species_list, done):
                    self._side.wait_event(ev)
                    E.finish_advance_p_all([sp])
                    if fixed:
                        self.exchange.boundary_p_fixed(self, [sp])
                    else:
                        self.exchange.boundary_p(self, species=[sp], check_empty=False)
                if fixed:
                    self.exchange.end_fixed(self, self.species_list)
            main.wait_stream(self._side)
            E.reduce_accumulator_array(aa)
            if fixed and not self.exchange.inject_may_emit:
                pass      # the later rounds are provably empty (no wall of this slab turns an injected particle into a mover)
            else:
                if fixed:
                    self.exchange.resolve(self)
                    E.finish_advance_p_all(self.species_list)
                for _ in range(self.num_comm_round - 1):
                    if not self.exchange.boundary_p(self):
                        break                     # no movers anywhere: the remaining rounds would be empty too
                for sp in self.species_list:
                    E.drop_unresolved_movers(sp, fa)          # advance.cc:78-101
            self.exchange._steps_seen += 1
        elif self.exchange is not None:
            E.finish_advance_p_all(self.species_list)
            E.reduce_accumulator_array(aa)
            for _ in range(self.num_comm_round):
                if not self.exchange.boundary_p(self):
                    break
            for sp in self.species_list:
                E.drop_unresolved_movers(sp, fa)              # advance.cc:78-101
        else:
            E.finish_advance_p_all(self.species_list)
            E.reduce_accumulator_array(aa)
            # one rank: the only movers are particles that hit an absorbing wall (boundary_p.cc:268-275)
            for sp in self.species_list:
                if sp.nm:
                    _, offs = E.boundary_pack(sp, [-1] * 6, fa)
                    o = offs.cpu()
                    if int(o[8] - o[7]) != 0:                     # class 7: neither absorbed nor sent anywhere
                        raise RuntimeError(f"species {sp.name}: particles left through a face that is neither local, "
                                           "absorbing nor shared with another rank (custom boundary handlers stay on "
                                           "the host)")
        fa.clear_jf()
        E.unload_accumulator_array(fa, aa)
        fa.synchronize_jf()
        if self.exchange is not None and self.overlap_exchange:
            # the shared-plane current sums travel on the side stream while the first half B advance (which does not
            # read jf) runs on the main one
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(priority=-1)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self.exchange.synchronize_jf(self)
            fa.advance_b(0.5)
            main.wait_stream(self._side)
        else:
            if self.exchange is not None:
                self.exchange.synchronize_jf(self)
            fa.advance_b(0.5)
        if self.exchange is not None:
            self.exchange.ghost_tang_b(self)
        fa.advance_e(1.0)
        fa.advance_b(0.5)
        self._maintain_fields(step)
        E.load_interpolator_array(ia, fa)
        self.g.g.step += 1

    def _allsum(self, values):
        return self.exchange.allsum(values) if self.exchange is not None else values

    def _maintain_fields(self, step):
        """advance.cc:138-176: Marder passes on div E and div B and the shared-face synchronisation."""
        fa, ex = self.field_array, self.exchange
        if self.clean_div_e_interval > 0 and step % self.clean_div_e_interval == 0:
            self.sync_counts()                 # accumulate_rho_p needs the particles this step's migration appended
            fa.clear_rhof()
            for sp in self.species_list:
                E.accumulate_rho_p(fa, sp)
            fa.synchronize_rho()
            if ex is not None:
                ex.synchronize_rho(self)
            for r in range(self.num_div_e_round):
                if ex is not None:
                    ex.ghost_norm_e(self)
                fa.compute_div_e_err()
                if r == 0 or r == self.num_div_e_round - 1:
                    err = fa.rms_finish(self._allsum(fa.rms_terms("vpb_compute_rms_div_e_err")))
                    self.cleaning_log.append((step, "div_e initial" if r == 0 else "div_e cleaned", err))
                fa.clean_div_e()
        if self.clean_div_b_interval > 0 and step % self.clean_div_b_interval == 0:
            for r in range(self.num_div_b_round):
                fa.compute_div_b_err()
                if r == 0 or r == self.num_div_b_round - 1:
                    err = fa.rms_finish(self._allsum(fa.rms_terms("vpb_compute_rms_div_b_err")))
                    self.cleaning_log.append((step, "div_b initial" if r == 0 else "div_b cleaned", err))
                if ex is not None:
                    ex.ghost_div_b(self)
                fa.clean_div_b()
        if self.sync_shared_interval > 0 and step % self.sync_shared_interval == 0:
            fa.synchronize_tang_e_norm_b(read=False)
            if ex is not None:
                ex.synchronize_tang_e_norm_b(self, fa._en)
            err = self._allsum([float(fa._en[0].item())])[0]
            self.cleaning_log.append((step, "desynchronization", err))

    def energies(self):
        """dump_energies row (src/vpic/dump.cc:38-77): field energies then one kinetic energy per species."""
        self.sync_counts()
        en_f = self.field_array.energy_f()
        en_p = [E.energy_p(sp, self.interpolator_array) for sp in self.species_list]
        return list(en_f) + en_p

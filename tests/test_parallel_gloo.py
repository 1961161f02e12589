"""N>1 host-side routing (vpic_b200/parallel.py NeighbourRing) over gloo on CPU, world_size 2 and 3.
What leaves a rank through its low face must arrive through the low neighbour's high face, also when both
neighbours are the same rank (world_size == 2), and empty messages must be skipped consistently on both sides.
SlabExchange.boundary_p is run the same way with the device kernels replaced by CPU stand-ins: per-species counts
and payloads, low-face-first injection order, and the rounds stopping after the first one in which nobody holds a
mover."""
import os
import socket
import sys
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vpic_b200.parallel import NeighbourRing
        ring = NeighbourRing(rank, world)
        ok = True
        # round 1: distinct payloads both ways
        out_lo = torch.full((5,), 10.0 * rank + 1)          # leaves through the low face
        out_hi = torch.full((7,), 10.0 * rank + 2)          # leaves through the high face
        in_lo, in_hi = torch.zeros(7), torch.zeros(5)
        ring.sendrecv(out_lo, out_hi, in_lo, in_hi)
        ok &= bool(torch.all(in_lo == 10.0 * ring.lo + 2))  # the low neighbour's high-face message
        ok &= bool(torch.all(in_hi == 10.0 * ring.hi + 1))  # the high neighbour's low-face message
        # round 2: counts first, then ragged payloads where some directions are empty (boundary_p pattern)
        n_lo, n_hi = (rank + 1) % 2 * 3, rank % 2 * 4       # even ranks send only low, odd ranks only high
        c_in_lo, c_in_hi = torch.zeros(1, dtype=torch.int32), torch.zeros(1, dtype=torch.int32)
        ring.sendrecv(torch.tensor([n_lo], dtype=torch.int32), torch.tensor([n_hi], dtype=torch.int32), c_in_lo, c_in_hi)
        exp_from_lo = ring.lo % 2 * 4
        exp_from_hi = (ring.hi + 1) % 2 * 3
        ok &= int(c_in_lo) == exp_from_lo and int(c_in_hi) == exp_from_hi
        in_lo, in_hi = torch.zeros(int(c_in_lo), 12), torch.zeros(int(c_in_hi), 12)
        ring.sendrecv(torch.full((n_lo, 12), float(rank)), torch.full((n_hi, 12), float(rank)), in_lo, in_hi)
        ok &= bool(torch.all(in_lo == float(ring.lo))) and bool(torch.all(in_hi == float(ring.hi)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_ring_routing(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res), res


def test_world_size_one_is_a_local_copy():
    sys.path.insert(0, ROOT)
    from vpic_b200.parallel import NeighbourRing
    ring = NeighbourRing(0, 1)
    a, b = torch.arange(4.0), torch.arange(4.0) + 10
    in_lo, in_hi = torch.zeros(4), torch.zeros(4)
    ring.sendrecv(a, b, in_lo, in_hi)
    assert torch.equal(in_hi, a) and torch.equal(in_lo, b)


def _migration_worker(rank, world, port, q):
    """SlabExchange.boundary_p with the device kernels replaced by CPU stand-ins: what is routed where, in which
    order it is injected, and when the rounds stop."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace as NS
        from vpic_b200 import engine as E, parallel
        cpu = torch.device("cpu")
        nv = 100
        grid = NS(range=[r * nv for r in range(world + 1)])
        dg = NS(rank=rank, world_size=world, g=grid, nx=4, ny=4, nz=4, device=cpu)
        ex = parallel.SlabExchange(dg, axis=1)

        def species(sid, n_lo, n_hi):
            # a mover record is 12 floats; tag it with (source rank, species, face, running index)
            rows = [[float(rank), float(sid), float(ex.f_lo), float(k)] + [0.0] * 8 for k in range(n_lo)] + \
                   [[float(rank), float(sid), float(ex.f_hi), float(k)] + [0.0] * 8 for k in range(n_hi)]
            return NS(name=f"s{sid}", nm=n_lo + n_hi, np=1000, max_np=10 ** 6, out=(n_lo, n_hi), rows=rows, got=[],
                      counters=torch.zeros(4, dtype=torch.int32))

        def fake_pack(sp, face_range, fa=None):
            assert face_range[ex.f_lo] == grid.range[ex.ring.lo] and face_range[ex.f_hi] == grid.range[ex.ring.hi]
            n_lo, n_hi = sp.out
            offs = torch.zeros(9, dtype=torch.int32)
            for c in range(9):                                  # classes ascending: faces 0..5, absorbed, no handler
                offs[c] = (n_lo if c > ex.f_lo else 0) + (n_hi if c > ex.f_hi else 0)
            inj = torch.tensor(sp.rows, dtype=torch.float32).reshape(-1, 12) if sp.rows else None
            sp.np -= sp.nm; sp.nm = 0; sp.out = (0, 0); sp.rows = []
            return inj, offs

        def fake_inject(sp, aa, ia, inj, n):
            assert inj.shape[0] == n
            sp.got += [tuple(int(x) for x in row[:4]) for row in inj.tolist()]
            sp.np += n

        E.boundary_pack, E.boundary_inject, E.finish_advance_p_all = fake_pack, fake_inject, (lambda sps: None)
        sps = [species(0, 2 + rank, 1), species(1, 0, 3 if rank % 2 == 0 else 0)]
        sim = NS(species_list=sps, field_array=None, accumulator_array=None, interpolator_array=None)
        first = ex.boundary_p(sim)                              # somebody holds movers: a full exchange
        second = ex.boundary_p(sim)                             # nobody does: one all-reduce, no exchange
        ok = first is True and second is False
        lo, hi = ex.ring.lo, ex.ring.hi
        for sid, sp in enumerate(sps):
            def sent(src, face):                                # what rank `src` sent out of `face` for this species
                n_lo, n_hi = (2 + src, 1) if sid == 0 else (0, 3 if src % 2 == 0 else 0)
                return [(src, sid, face, k) for k in range(n_lo if face == ex.f_lo else n_hi)]
            # injection order of the reference: low face first.  Through my low face arrive the low neighbour's
            # high-face movers, through my high face the high neighbour's low-face movers.
            ok &= sp.got == sent(lo, ex.f_hi) + sent(hi, ex.f_lo)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_migration_routing_and_round_termination(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_migration_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res), res


def _fixed_worker(rank, world, port, q):
    """SlabExchange.boundary_p_fixed (fixed-capacity messages, counts in the headers, deferred bookkeeping) with the
    device kernels replaced by CPU stand-ins."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace as NS
        from vpic_b200 import engine as E, parallel
        cpu = torch.device("cpu")
        nv = 100
        grid = NS(range=[r * nv for r in range(world + 1)])
        dg = NS(rank=rank, world_size=world, g=grid, nx=4, ny=4, nz=4, device=cpu)
        ex = parallel.SlabExchange(dg, axis=1)
        assert ex.inject_may_emit is False

        def n_out(src, sid, step):
            return ((2 + src + step, 1) if sid == 0 else (0, 3 if src % 2 == 0 else 0))

        def species(sid):
            return NS(name=f"s{sid}", id=sid, nm=0, np=1000, max_np=10 ** 6, rows=[], out=(0, 0), placed={},
                      counters=torch.zeros(4, dtype=torch.int32))

        def load(sp, step):
            n_lo, n_hi = n_out(rank, sp.id, step)
            sp.rows = [[float(rank), float(sp.id), float(ex.f_lo), float(k)] + [0.0] * 8 for k in range(n_lo)] + \
                      [[float(rank), float(sp.id), float(ex.f_hi), float(k)] + [0.0] * 8 for k in range(n_hi)]
            sp.nm, sp.out = n_lo + n_hi, (n_lo, n_hi)

        def fake_pack(sp, face_range, fa=None, absorb_all=False):
            n_lo, n_hi = sp.out
            offs = torch.zeros(9, dtype=torch.int32)
            for c in range(9):
                offs[c] = (n_lo if c > ex.f_lo else 0) + (n_hi if c > ex.f_hi else 0)
            inj = torch.tensor(sp.rows, dtype=torch.float32).reshape(-1, 12) if sp.rows else None
            sp.np -= sp.nm; sp.nm = 0; sp.out = (0, 0); sp.rows = []
            return inj, offs

        def fake_stage(inj, offs, face, cap, sp_id, msg, status):
            first, n = int(offs[face]), int(offs[face + 1] - offs[face])
            msg.view(torch.int32)[0:4] = torch.tensor([n, sp_id, cap, 0], dtype=torch.int32)
            if n > cap:
                status[0] |= 1
            status[1] = max(int(status[1]), n)
            if n:
                msg[4:4 + 12 * min(n, cap)] = inj[first:first + min(n, cap)].reshape(-1)

        def fake_inject_msg(sp, aa, ia, msg, cap, added, status):
            n = min(int(msg.view(torch.int32)[0]), cap)
            rec = msg[4:4 + 12 * n].reshape(n, 12)
            base = sp.np + int(added[0])
            for j in range(n):                                   # the reference appends the LAST record first
                sp.placed[base + j] = tuple(int(x) for x in rec[n - 1 - j, :4])
            added[0] += n

        E.boundary_pack, E.boundary_stage, E.boundary_inject_msg = fake_pack, fake_stage, fake_inject_msg
        sps = [species(0), species(1)]
        sim = NS(species_list=sps, field_array=None, accumulator_array=None, interpolator_array=None)
        for sp in sps:
            ex._alloc(sp, 8)                                     # rounded up to the minimum capacity
        ok = True
        lo, hi = ex.ring.lo, ex.ring.hi
        for step in range(2):
            for sp in sps:
                load(sp, step)
                sp.placed = {}
            np_before = [sp.np for sp in sps]
            for sp in sps:
                ex.boundary_p_fixed(sim, [sp])
            ex.end_fixed(sim, sps)
            ok &= all(sp.np == np_before[k] - sum(n_out(rank, k, step)) for k, sp in enumerate(sps))   # deferred
            ex.resolve(sim)
            for k, sp in enumerate(sps):
                from_lo = [(lo, k, ex.f_hi, i) for i in range(n_out(lo, k, step)[1])]
                from_hi = [(hi, k, ex.f_lo, i) for i in range(n_out(hi, k, step)[0])]
                base = np_before[k] - sum(n_out(rank, k, step))
                want = {base + j: r for j, r in enumerate(from_lo[::-1] + from_hi[::-1])}
                ok &= sp.placed == want
                ok &= sp.np == base + len(from_lo) + len(from_hi)
        # a message larger than its capacity is reported at the next resolve, never dropped silently
        ex._mig[0]["cap"] = 1
        load(sps[0], 0)
        ex.boundary_p_fixed(sim, [sps[0]])
        ex.end_fixed(sim, sps)
        try:
            ex.resolve(sim)
            ok = False
        except RuntimeError as e:
            ok &= "exceeded its capacity" in str(e)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_fixed_capacity_migration(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fixed_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res), res

"""N>1 host-side routing (vpic_b200/parallel.py NeighbourRing) over gloo on CPU, world_size 2 and 3.
What leaves a rank through its low face must arrive through the low neighbour's high face, also when both
neighbours are the same rank (world_size == 2), and empty messages must be skipped consistently on both sides."""
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

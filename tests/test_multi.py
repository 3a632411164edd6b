"""CPU tests of the multi-GPU host logic (no GPU): the tile schedule of mecat_b200/multi.py covers
every (index volume, query volume >= index volume, query read) exactly once, and the ring rotation
delivers every block to every rank -- exercised with real processes over gloo (world size 2 and 3)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mecat_b200 import multi  # noqa: E402


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
def test_schedule_covers_every_tile_once(world):
    reads = [1000 + 37 * v for v in range(world)]
    cover = {}
    load = []
    for rank in range(world):
        units = 0
        for step in range(world):
            for s, v, rb, re in multi.tile_work(world, rank, step, reads):
                assert v >= s and 0 <= rb < re <= reads[v]
                for r in (rb, re - 1):
                    pass
                key = (s, v)
                cover.setdefault(key, []).append((rb, re))
                units += (re - rb) / reads[v]
        load.append(units)
    assert set(cover) == {(s, v) for s in range(world) for v in range(s, world)}
    for (s, v), parts in cover.items():
        parts.sort()
        assert parts[0][0] == 0 and parts[-1][1] == reads[v]
        for a, b in zip(parts, parts[1:]):
            assert a[1] == b[0]
    # mirror pairing balances the triangular loop: every rank does (N+1)/2 tile equivalents
    assert max(load) - min(load) < 1e-9 + 0.01 and abs(sum(load) - world * (world + 1) / 2) < 1e-6


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reads = [100 + v for v in range(world)]

    class Blk:
        def __init__(self, v):
            self.t = torch.full((4,), float(v))
            self.v = v

    own, a, b = Blk(rank), Blk(-1), Blk(-1)
    nxt, prv = multi.ring_neighbours(world, rank)
    seen = []

    class H:
        def __init__(self, reqs, blk, v):
            self.reqs, self.blk, self.v = reqs, blk, v

        def wait(self):
            for r in self.reqs:
                r.wait()
            self.blk.v = self.v

    def exchange(cur, sp):
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, cur.t, nxt), dist.P2POp(dist.irecv, sp.t, prv)])
        return H(reqs, sp, (cur.v - 1) % world)

    def compute(step, blk):
        assert int(blk.t[0].item()) == blk.v == multi.block_at(world, rank, step)   # payload really is that block
        seen.append((blk.v, multi.tile_work(world, rank, step, reads)))

    multi.run_ring(world, rank, own, a, b, exchange, compute)
    assert int(own.t[0].item()) == rank            # the rank's own block is never overwritten
    gathered = [None] * world
    dist.all_gather_object(gathered, seen)
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_rotation_over_gloo(world):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tiles = set()
    for rank, seen in enumerate(gathered):
        assert sorted(v for v, _ in seen) == list(range(world))       # every block visited every rank
        for _, items in seen:
            for s, v, rb, re in items:
                tiles.add((s, v, rb, re))
    pairs = {(s, v) for s, v, _, _ in tiles}
    assert pairs == {(s, v) for s in range(world) for v in range(s, world)}


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_read_and_code_slices_partition(world):
    n = 100003
    for diag in (True, False):
        sl = multi.read_slices(world, n, diagonal=diag)
        assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        assert all(b > a for a, b in sl)
    cs = multi.code_slices(world)
    assert cs[0][0] == 0 and cs[-1][1] == 1 << 26 and all(a[1] == b[0] for a, b in zip(cs, cs[1:]))
    assert all(lo % 256 == 0 and hi % 256 == 0 for lo, hi in cs)


def _exchange_worker(rank, world, port, mode, out):
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # uneven slices, one of them empty when there are three ranks
    sizes = [1000, 37, 0][:world] if world == 3 else [129, 1000]
    bounds = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    full = torch.arange(int(bounds[-1]), dtype=torch.int32) * 7 + 3           # what every rank must end up with
    pos = torch.full_like(full, -1)
    pos[int(bounds[rank]):int(bounds[rank + 1])] = full[int(bounds[rank]):int(bounds[rank + 1])]   # only the own slice is known
    multi.exchange_slices(dist, pos, bounds, rank, world, mode)
    ok = bool(torch.equal(pos, full))
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        out.put(flags)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "allgather"), (3, "allgather"), (2, "broadcast")])
def test_index_slice_exchange_over_gloo(world, mode):
    """Strong scaling: every rank builds one code slice of the k-mer positions; after the exchange all ranks hold all
    slices (padded all-gather + compaction, or the per-slice broadcasts it replaced)."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    flags = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert flags == [True] * world

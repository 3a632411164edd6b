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


# ---------------------------------------------------------------------------------------------- bench drivers on CPU
class _Args:
    def __init__(self, **kw):
        self.reads, self.steps, self.warmup, self.volumes, self.ring_reads = 0, 2, 1, 0, 0
        self.__dict__.update(kw)


class _StubCtx:
    """Stands in for mecat_b200.Context in the CPU runs of the two bench drivers: a 'device volume' remembers what it was
    built from, an 'index' is real tensors over a small code space (so the histogram all-gather and the position exchange
    move real data), a tile call returns one record per query read of its range and logs what it was asked to do."""
    NCODES = 1 << 12

    def __init__(self, torch):
        self.torch = torch
        self.calls = []
        self.closed = False

    # volumes
    def volume_from_device(self, nr, nb, sid, osz, pac):
        assert osz.shape == (nr, 2)
        return {"nr": nr, "nb": nb, "sid": sid, "tag": int(pac[0].item()), "sum": int(pac.to(self.torch.int64).sum().item())}

    def release_volume(self, d):
        d["released"] = True

    # index in two stages (strong mode)
    def index_count_part(self, dvol, lo, hi):
        t = self.torch
        counts = t.zeros(self.NCODES, dtype=t.int32)
        codes = t.arange(lo, hi, dtype=t.int64)
        counts[lo:hi] = ((codes * 7 + 3) % 5).to(t.int32)
        return {"counts": counts, "begin": None, "pos": None, "vol": dvol}

    def index_device_arrays(self, idx):
        return idx["counts"], idx["begin"], idx["pos"], (0 if idx["pos"] is None else idx["pos"].numel())

    def index_finish_part(self, dvol, idx, lo, hi):
        t = self.torch
        c = idx["counts"].to(t.int64)
        begin = t.zeros(self.NCODES + 1, dtype=t.int64)
        begin[1:] = t.cumsum(c, 0)
        idx["begin"] = begin.to(t.int32)
        pos = t.full((int(begin[-1].item()),), -1, dtype=t.int32)
        a, b = int(begin[lo].item()), int(begin[hi].item())
        pos[a:b] = t.arange(a, b, dtype=t.int32) * 3 + 1
        idx["pos"] = pos

    def index_build(self, dvol):
        return {"vol": dvol}

    def release_index(self, idx):
        idx["released"] = True

    def pw_tile_range(self, idx, dref, dq, params, rb, re):
        t = self.torch
        if idx.get("pos") is not None:
            want = t.arange(idx["pos"].numel(), dtype=t.int32) * 3 + 1
            assert t.equal(idx["pos"], want), "a rank seeds against an incomplete index"
        assert not dref.get("released") and not dq.get("released") and not idx.get("released")
        self.calls.append((dref["sid"], dq["sid"], dq["tag"], rb, re))
        return range(re - rb)

    def stats(self):
        names = ["index_count", "index_scan", "index_fill", "index_sort", "seed", "walk", "extend"]
        return {"kernel_ms": dict.fromkeys(names, 1.0), "kernel_launches": dict.fromkeys(names, 1), "gpu_launches": 7,
                "h2d_bytes": 0, "d2h_bytes": 0}

    def reset_stats(self):
        pass

    def close(self):
        self.closed = True


class _NoClocks:
    def __init__(self, *a):
        pass

    def stop(self):
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}


def _host_volume(v, num_reads):
    import numpy as np

    class HV:
        pass
    h = HV()
    lens = 40 + (np.arange(num_reads) % 7)
    h.num_reads = num_reads
    h.num_bases = int(lens.sum()) + num_reads
    off = np.concatenate([[0], np.cumsum(lens + 1)[:-1]])
    h.offset_size = np.stack([off, lens], 1).astype(np.int32)
    h.pac = np.full((h.num_bases + 3) // 4, v + 1, dtype=np.uint8)       # payload = volume number + 1
    h.start_read_id = v * 1000
    return h


def _bench_worker(rank, world, port, mode, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _StubCtx(torch)
    env = multi.Env(dist, torch, rank, world, "cpu", ctx, ncodes=_StubCtx.NCODES, pin=False)
    common = ("metric", "unit", lambda n: {"workload": "stub"}, None, None, _NoClocks, None, lambda *a, **k: {"stub": True})
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        if mode == "strong":
            line = multi.run_bench_strong(_Args(), *common, env=env, volume=_host_volume(0, 101))
        else:
            V = 4
            sets = multi.volume_sets(world, V)
            hv = {v: _host_volume(v, 60 + v) for v in sets[rank]}
            line = multi.run_bench(_Args(volumes=V), *common, env=env, host_volumes=hv)
    out.put((rank, line, ctx.calls, ctx.closed, buf.getvalue()))


@pytest.mark.parametrize("world,mode", [(2, "strong"), (2, "ring"), (1, "ring"), (4, "ring")])
def test_bench_drivers_run_on_cpu(world, mode):
    """Both multi-GPU bench drivers, end to end over gloo with a stub context: warm-up, timed steps, the end-to-end steps
    and the JSON line (a NameError in either would have shipped in round 1).  Strong: every rank seeds against a complete
    index and the ranks' read slices partition the volume.  Ring: every (index volume <= query volume) tile is covered
    exactly once per step, with the right payload under every volume handle."""
    import json
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_bench_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    line = got[0][1]
    assert line is not None and all(g[1] is None for g in got[1:]) and all(g[3] for g in got)
    assert json.loads(got[0][4].strip().splitlines()[-1])["n_gpus"] == world      # rank 0 printed exactly the line it returned
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "e2e",
              "gpu_launches", "roofline", "cpu_baseline", "config"):
        assert k in line, k
    nsteps = 1 + 2 + 2                                        # warm-up + timed + end-to-end steps
    calls = [c for g in got for c in g[2]]
    if mode == "strong":
        assert line["pairs_per_step"] == 101 and line["scaling"] == "strong"
        spans = sorted((rb, re) for _, _, _, rb, re in calls)
        per_step = spans[::nsteps]
        assert per_step[0][0] == 0 and per_step[-1][1] == 101 and all(a[1] == b[0] for a, b in zip(per_step, per_step[1:]))
        assert line["e2e"]["h2d_bytes_per_step"] < 2 * (len(_host_volume(0, 101).pac) + 16 * world + 101 * 8)     # one upload of the volume, not N
    else:
        V = 4
        cover = {}
        for sid_s, sid_v, tag, rb, re in calls:
            s_, v_ = sid_s // 1000, sid_v // 1000
            assert tag == v_ + 1                              # the packed bytes under the handle are that volume's
            cover.setdefault((s_, v_), []).append((rb, re))
        assert set(cover) == {(s_, v_) for s_ in range(V) for v_ in range(s_, V)}
        for (s_, v_), parts in cover.items():
            parts.sort()
            assert len(parts) % nsteps == 0
            one = parts[::nsteps]
            assert one[0][0] == 0 and one[-1][1] == 60 + v_ and all(a[1] == b[0] for a, b in zip(one, one[1:]))
        assert line["pairs_per_step"] == sum((60 + v_) * (v_ + 1) for v_ in range(V))
        assert line["config"]["tiles"] == 10 and line["scaling"] == "strong"


@pytest.mark.parametrize("world,volumes", [(1, 8), (2, 8), (4, 8), (8, 8), (2, 4), (3, 6)])
def test_ring_work_is_balanced_and_complete(world, volumes):
    reads = [1000 + 10 * v for v in range(volumes)]
    cover, load = {}, []
    for rank in range(world):
        units = 0.0
        for step in range(world):
            for s, v, rb, re in multi.ring_work(world, rank, step, volumes, reads):
                assert s <= v and 0 <= rb < re <= reads[v]
                cover.setdefault((s, v), []).append((rb, re))
                units += (re - rb) / reads[v]
        load.append(units)
    assert set(cover) == {(s, v) for s in range(volumes) for v in range(s, volumes)}
    for parts in cover.values():
        parts.sort()
        assert parts[0][0] == 0 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    # mirror pairing: every rank serves the same number of tiles (diagonal tiles are cut off-centre on purpose)
    tiles = volumes * (volumes + 1) / 2
    assert max(load) - min(load) <= 0.45 * 2 * volumes / world and abs(sum(load) - tiles) < 1e-6

"""Multi-GPU tiling of the all-vs-all overlap job: one process per GPU (torch.distributed).

The reference runs the upper-triangular loop over (index volume s, query volume v >= s) serially
(src/mecat2pw/pw.cpp:65-81, pw_impl.cpp:859-879).  Here, with one volume per rank:

  * rank g keeps the k-mer indices of volumes g and N-1-g (its "mirror"); the two ranks of a
    pair split every visiting query volume's reads in half, so each rank does (N+1)/2 tile
    equivalents instead of the 1..N of a plain triangular assignment;
  * packed query volumes (~0.4 GB each) travel round the ring of ranks with NCCL send/recv over
    NVLink, double buffered so the transfer of the next block overlaps the compute of the
    current one.  That rotation is the only collective of the path: the payload is input data,
    nothing is reduced, so there is no compute step to fuse it with.

The schedule is plain Python and backend agnostic (gloo on CPU in the tests, NCCL on GPUs).
"""
import json
import os
import sys
import time

import numpy as np


def served_indices(world, rank):
    """Index volumes whose tiles this rank works on."""
    mirror = world - 1 - rank
    return [rank] if mirror == rank else sorted((rank, mirror))


def read_share(world, rank, num_reads):
    """[begin, end) of a visiting block's reads this rank handles (the pair splits them in half)."""
    mirror = world - 1 - rank
    if mirror == rank:
        return 0, num_reads
    half = num_reads // 2
    return (0, half) if rank < mirror else (half, num_reads)


def block_at(world, rank, step):
    """Query block resident on `rank` at ring step `step` (blocks move to rank+1 every step)."""
    return (rank - step) % world


def tile_work(world, rank, step, reads_in_block):
    """Work items (index volume, query block, read_begin, read_end) of one ring step."""
    v = block_at(world, rank, step)
    rb, re = read_share(world, rank, reads_in_block[v])
    return [(s, v, rb, re) for s in served_indices(world, rank) if v >= s and re > rb]


def ring_neighbours(world, rank):
    return (rank + 1) % world, (rank - 1) % world


def run_ring(world, rank, own_block, buf_a, buf_b, exchange, compute):
    """Generic rotation: `compute(step, block)` on the resident block while `exchange(send, recv_buf)`
    moves blocks one rank up (double buffered; the rank's own block is never overwritten).
    exchange returns a handle with .wait()."""
    cur = own_block
    spare, other = buf_a, buf_b
    for step in range(world):
        pending = None
        if step + 1 < world:
            pending = exchange(cur, spare)
        compute(step, cur)
        if pending is not None:
            pending.wait()
            cur, spare, other = spare, other, (cur if cur is not own_block else other)
            if spare is cur:
                spare = other


def code_slices(world, n=1 << 26):
    """Equal slices of the 2^26 k-mer codes, aligned to 256 codes (the index kernels' granularity)."""
    cuts = [((n * r // world) // 256) * 256 for r in range(world)] + [n]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def read_slices(world, num_reads, diagonal=True, seed_w=131.0, extend_w=417.0):
    """Split points of a tile's query reads.  In a diagonal tile (query volume == index volume) read q
    can only pair with reads <= q (pw_impl.cpp:370), so extension work grows linearly with the read
    ordinal while seeding work is flat: cost(x) = seed_w * x + extend_w * x^2 for the first fraction x
    of the reads (weights = measured kernel milliseconds of the two phases on configs[1]).  Off
    diagonal tiles are uniform."""
    if not diagonal or world == 1:
        return [(num_reads * r // world, num_reads * (r + 1) // world) for r in range(world)]
    a, b = seed_w, extend_w
    cuts = []
    for r in range(world + 1):
        t = (a + b) * r / world
        x = (-a + (a * a + 4.0 * b * t) ** 0.5) / (2.0 * b)
        cuts.append(min(num_reads, int(round(x * num_reads))))
    cuts[0], cuts[-1] = 0, num_reads
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def exchange_slices(dist, pos, bounds, rank, world, mode="allgather"):
    """Strong-scaling index exchange: rank q owns pos[bounds[q]:bounds[q+1]] (its code slice of the k-mer positions);
    afterwards every rank holds all of pos.  "allgather": ONE padded all-gather (every NVLink port busy at once) followed
    by a local compaction of the foreign slices; "broadcast": one broadcast per slice (the first implementation).
    Works on any backend / device the tensor lives on (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    sizes = [int(bounds[q + 1] - bounds[q]) for q in range(world)]
    if mode == "allgather" and max(sizes) > 0:
        m = (max(sizes) + 63) // 64 * 64
        send = torch.empty(m, dtype=pos.dtype, device=pos.device)
        send[:sizes[rank]] = pos[int(bounds[rank]):int(bounds[rank + 1])]
        recv = torch.empty(world * m, dtype=pos.dtype, device=pos.device)
        dist.all_gather_into_tensor(recv, send)
        for q in range(world):
            if q != rank and sizes[q]:
                pos[int(bounds[q]):int(bounds[q + 1])] = recv[q * m:q * m + sizes[q]]
    else:
        reqs = [dist.broadcast(pos[int(bounds[q]):int(bounds[q + 1])], src=q, async_op=True) for q in range(world) if sizes[q] > 0]
        for r in reqs:
            r.wait()


class _DevMem:
    """Raw device memory as a __cuda_array_interface__ object (zero-copy torch view of library memory)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_view(ptr, count, torch_dtype, itemsize):
    import torch
    return torch.as_tensor(_DevMem(ptr, count * itemsize), device="cuda").view(torch_dtype)


# ------------------------------------------------------------------------------------------ bench plumbing
class Env:
    """What the two bench drivers below need from the machine: the process group, the device, a context of the library.
    On a GPU box that is NCCL + cuda:<local rank> + mecat_b200.Context; the CPU suite injects gloo + "cpu" + a stub
    context (tests/test_multi.py), so the drivers' control flow runs in the `-m "not gpu"` tests as well."""

    def __init__(self, dist, torch, rank, world, device, ctx, ncodes=1 << 26, pin=True):
        self.dist, self.torch, self.rank, self.world, self.device, self.ctx = dist, torch, rank, world, device, ctx
        self.ncodes, self.pin = ncodes, pin
        self.cuda = str(device).startswith("cuda")

    def sync(self):
        if self.cuda:
            self.torch.cuda.synchronize()

    def view(self, ptr, count, dtype, itemsize):
        """Library device memory as a tensor (zero copy); the stub context hands out tensors directly."""
        if not self.cuda:
            return ptr[:count]
        return device_view(ptr, count, dtype, itemsize)

    def pinned(self, t):
        return t.pin_memory() if (self.cuda and self.pin) else t

    def log(self, *a):
        print("[bench r%d]" % self.rank, *a, file=sys.stderr, flush=True)

    def timed(self, nsteps, one_step, *args):
        """Barrier + device sync on both sides, MAX over ranks of the wall time, SUM over ranks of the records."""
        torch, dist = self.torch, self.dist
        dist.barrier(); self.sync()
        t0 = time.perf_counter()
        n = 0
        for _ in range(nsteps):
            n += one_step(*args)
        self.sync(); dist.barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([n], dtype=torch.int64, device=self.device)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        return int(c.item()), float(t.item())


def gpu_env():
    import torch
    import torch.distributed as dist
    import mecat_b200
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    env = Env(dist, torch, rank, world, dev, mecat_b200.Context(local))
    env.local = local
    return env


def measured_peaks():
    try:
        return json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def byte_slices(nbytes, world):
    """Equal 16-byte aligned slices of the packed volume: every rank uploads one, NVLink carries the rest."""
    chunk = ((nbytes + world - 1) // world + 15) // 16 * 16
    return chunk, [(min(nbytes, r * chunk), min(nbytes, (r + 1) * chunk)) for r in range(world)]


# ------------------------------------------------------------------------------------------ benchmark (strong scaling)
def run_bench_strong(args, METRIC, UNIT, workload_config, make_reads, tmp_root, ClockSampler, cpu_sample, roofline_for,
                     env=None, volume=None):
    """N GPUs share ONE tile (BASELINE configs[1]): every rank builds the slice [code_lo, code_hi) of the k-mer index,
    the slices are exchanged over NCCL (all-gather of the histogram, padded all-gather of the position slices), and rank r
    seeds / extends its slice of the query reads.  End to end every rank uploads 1/N of the packed volume over PCIe and
    the ranks all-gather it over NVLink (one 0.4 GB input, not N uploads of it).
    `env` / `volume`: injected by the CPU tests (gloo, stub context, small volume)."""
    import mecat_b200
    env = env or gpu_env()
    torch, dist, ctx, rank, world, dev = env.torch, env.dist, env.ctx, env.rank, env.world, env.device
    log = env.log

    READS, GENOME, SEED = 100000, 100000000, 11
    if args.reads:
        READS = args.reads; GENOME = args.reads * 1000
    if volume is None:
        d = tmp_root()
        fa = os.path.join(d, "reads_%d_%d.fa" % (READS, SEED))
        wrk = os.path.join(d, "wrk_%d" % READS)
        if rank == 0:
            make_reads(fa, READS, GENOME, SEED)
            mecat_b200.split_dataset(fa, wrk)
        dist.barrier()
        volume = mecat_b200.HostVolume.load(os.path.join(wrk, "vol0"))
    vol = volume
    nbytes = len(vol.pac)
    chunk, bsl = byte_slices(nbytes, world)
    pac = env.pinned(torch.zeros(chunk * world, dtype=torch.uint8))
    pac.numpy()[:nbytes] = vol.pac
    osz = env.pinned(torch.from_numpy(vol.offset_size.reshape(-1).copy()))
    osz_np = vol.offset_size.reshape(-1, 2)
    params = mecat_b200.pw_params(task=1)
    lo, hi = code_slices(world, env.ncodes)[rank]
    rb, re = read_slices(world, vol.num_reads)[rank]
    cuts = torch.tensor([c[0] for c in code_slices(world, env.ncodes)] + [env.ncodes], dtype=torch.int64, device=dev)
    resident = [None]
    pac_d = torch.empty(chunk * world, dtype=torch.uint8, device=dev)
    # index exchange: "allgather" = one padded all-gather of the position slices (all NVLink ports busy at once) followed by
    # a device-side compaction; "broadcast" = one broadcast per slice (the first implementation, kept for comparison)
    exchange = os.environ.get("MECAT_STRONG_EXCHANGE", "allgather")
    phase = {"upload": 0.0, "count": 0.0, "counts_allgather": 0.0, "finish": 0.0, "positions_exchange": 0.0, "tile": 0.0}
    io = {"h2d": 0, "nvlink": 0}

    def lap(name, t0):
        env.sync()
        t1 = time.perf_counter()
        phase[name] += (t1 - t0) * 1e3
        return t1

    def upload():
        # own byte slice over PCIe, the other N - 1 over NVLink
        a, b = bsl[rank]
        mine = pac_d[rank * chunk:(rank + 1) * chunk]
        mine[:b - a].copy_(pac[a:b], non_blocking=True)
        io["h2d"] += (b - a) + osz.numel() * 4
        if world > 1:
            dist.all_gather_into_tensor(pac_d, mine.clone())
            io["nvlink"] += (world - 1) * chunk
        env.sync()
        return ctx.volume_from_device(vol.num_reads, vol.num_bases, 0, osz_np, pac_d.data_ptr() if env.cuda else pac_d)

    def one_step(e2e):
        t0 = time.perf_counter()
        if e2e or resident[0] is None:
            if resident[0] is not None:
                ctx.release_volume(resident[0])
            resident[0] = upload()
        dvol = resident[0]
        t0 = lap("upload", t0)
        idx = ctx.index_count_part(dvol, lo, hi)
        t0 = lap("count", t0)
        cptr, bptr, _, _ = ctx.index_device_arrays(idx)
        counts = env.view(cptr, env.ncodes, torch.int32, 4)
        if world > 1:
            parts = [counts[a:b] for a, b in code_slices(world, env.ncodes)]
            if len({p.numel() for p in parts}) == 1:
                dist.all_gather_into_tensor(counts, parts[rank].clone())
            else:
                for q in range(world):
                    dist.broadcast(parts[q], src=q)
            env.sync()
        t0 = lap("counts_allgather", t0)
        ctx.index_finish_part(dvol, idx, lo, hi)
        t0 = lap("finish", t0)
        if world > 1:
            _, bptr, pptr, nk = ctx.index_device_arrays(idx)
            begin = env.view(bptr, env.ncodes + 1, torch.int32, 4)
            bounds = begin[cuts].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            pos = env.view(pptr, nk, torch.int32, 4)
            exchange_slices(dist, pos, bounds, rank, world, exchange)
            env.sync()
        t0 = lap("positions_exchange", t0)
        rec = ctx.pw_tile_range(idx, dvol, dvol, params, rb, re)
        ctx.release_index(idx)
        lap("tile", t0)
        return len(rec)

    for i in range(args.warmup):
        n, dt = env.timed(1, one_step, False)
        if rank == 0:
            log("warmup %d: %d pairs in %.3f s" % (i, n, dt))
    ctx.reset_stats()
    for k in phase:
        phase[k] = 0.0
    sampler = ClockSampler(getattr(env, "local", 0)) if rank == 0 else None
    pairs, dt = env.timed(args.steps, one_step, False)
    stats = ctx.stats()
    phase_ms = {k: round(v / args.steps, 3) for k, v in phase.items()}
    clocks = sampler.stop() if sampler else None
    esteps = max(1, min(args.steps, 3))
    ctx.reset_stats()
    io["h2d"] = io["nvlink"] = 0
    for k in phase:
        phase[k] = 0.0
    epairs, edt = env.timed(esteps, one_step, True)
    estats = ctx.stats()
    ephase_ms = {k: round(v / esteps, 3) for k, v in phase.items()}
    iot = torch.tensor([estats["h2d_bytes"] + io["h2d"], estats["d2h_bytes"], io["nvlink"]], dtype=torch.int64, device=dev)
    dist.all_reduce(iot, op=dist.ReduceOp.SUM)
    line = None
    if rank == 0:
        peaks = measured_peaks()
        cfg = workload_config(1)
        cfg["parallelism"] = ("%d gpus sharing one tile: k-mer index built in %d code slices exchanged over NCCL "
                              "(all-gather), query reads split %d ways; end to end each rank uploads 1/%d of the packed "
                              "volume and the ranks all-gather it over NVLink" % (world, world, world, world))
        if args.reads:
            cfg = dict(cfg, reads=READS, genome=GENOME, workload="REDUCED debug workload (%d reads)" % READS)
        line = {
            "metric": METRIC, "value": pairs / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": epairs / edt, "unit": UNIT, "h2d_bytes_per_step": int(iot[0].item()) // esteps,
                    "d2h_bytes_per_step": int(iot[1].item()) // esteps, "nvlink_bytes_per_step": int(iot[2].item()) // esteps,
                    "ms_per_step": 1000.0 * edt / esteps, "steps": esteps, "phase_ms_per_step_rank0": ephase_ms},
            "gpu_launches": stats["gpu_launches"] * world,
            "roofline": roofline_for(stats, peaks, args.steps, world=world, clocks=clocks),
            "cpu_baseline": {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                             "sample": "measured at N=1 only (bench.py --gpus 1)"},
            "pairs_per_step": pairs // args.steps,
            "phase_ms_per_step_rank0": phase_ms, "index_exchange": exchange,
            "kernel_ms_per_step_rank0": {k: round(v / args.steps, 3) for k, v in stats["kernel_ms"].items()},
        }
        print(json.dumps(line))
    if resident[0] is not None:
        ctx.release_volume(resident[0])
    ctx.close()
    dist.destroy_process_group()
    return line


# ------------------------------------------------------------------------------------------ benchmark (block rotation)
def volume_sets(world, volumes):
    """Volumes owned by each rank (consecutive, volumes % world == 0)."""
    p = volumes // world
    return [list(range(r * p, (r + 1) * p)) for r in range(world)]


def ring_work(world, rank, step, volumes, reads_in_volume):
    """Work items (index volume s, query volume v, read_begin, read_end) of one ring step of the V-volume job: the set of
    volumes resident on `rank` at `step` is the one `rank - step` owns; the rank serves the index volumes of its own set
    and of its mirror rank's set, and the two ranks of a pair split every visiting volume's reads (a diagonal tile at
    the point that balances its triangular extension work, read_slices)."""
    sets = volume_sets(world, volumes)
    mirror = world - 1 - rank
    served = sorted(set(sets[rank]) | set(sets[mirror]))
    items = []
    for v in sets[block_at(world, rank, step)]:
        for s in served:
            if s > v:
                continue
            n = reads_in_volume[v]
            if mirror == rank:
                rb, re = 0, n
            else:
                cut = read_slices(2, n, diagonal=(s == v))[0][1]
                rb, re = (0, cut) if rank < mirror else (cut, n)
            if re > rb:
                items.append((s, v, rb, re))
    return items


def run_bench(args, METRIC, UNIT, workload_config, make_reads, tmp_root, ClockSampler, cpu_sample, roofline_for,
              env=None, host_volumes=None):
    """BASELINE configs[4] as a strong-scaling job: V volumes (default 8 x 125 000 reads = 1 M reads over a 1 Gb genome,
    36 tiles), the same at every N that divides V.  Rank g owns V/N volumes, keeps the k-mer indices of its own and of
    rank N-1-g's volumes, and the packed volume sets rotate round the ring of ranks over NCCL send/recv (double buffered
    behind the compute).  N = 1 runs all tiles on one device -- the baseline of the 1 -> 8 scaling figure.
    `env` / `host_volumes`: injected by the CPU tests (gloo, stub context, small volumes)."""
    import mecat_b200
    env = env or gpu_env()
    torch, dist, ctx, rank, world, dev = env.torch, env.dist, env.ctx, env.rank, env.world, env.device
    log = env.log

    V = args.volumes or 8
    if V % world:
        raise SystemExit("--mode ring needs a number of volumes (%d) that the number of GPUs (%d) divides" % (V, world))
    READS, SEED = args.ring_reads or 125000, 11
    if args.reads:
        READS = args.reads
    GENOME = READS * 1000                                    # 15x of 15 kb reads, per volume
    sets = volume_sets(world, V)
    own_vols = sets[rank]
    p = len(own_vols)
    if host_volumes is None:
        d = tmp_root()
        exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "gen_reads")
        host_volumes = {}
        import subprocess
        for v in own_vols:
            fa = os.path.join(d, "reads_V%d_v%d_%d_%d.fa" % (V, v, READS, SEED))
            wrk = os.path.join(d, "wrk_V%d_v%d_%d" % (V, v, READS))
            if not os.path.exists(os.path.join(wrk, "vol0")):
                # reads [v*READS, (v+1)*READS) of the V*READS-read data set over the V*GENOME genome: the same data at every N
                subprocess.check_call([exe, fa + ".tmp", str(READS), str(GENOME * V), str(SEED), "15000", "1500", "0.15", "-",
                                       str(v * READS)])
                os.replace(fa + ".tmp", fa)
                names = mecat_b200.split_dataset(fa, wrk)
                assert len(names) == 1, "a ring volume must fit one volume file"
                os.remove(fa)
            hv = mecat_b200.HostVolume.load(os.path.join(wrk, "vol0"))
            hv.start_read_id = v * READS
            host_volumes[v] = hv
    # pinned host copies (source of the H2D inside the e2e region)
    pac_h, osz_h = {}, {}
    for v in own_vols:
        hv = host_volumes[v]
        t = env.pinned(torch.zeros((len(hv.pac) + 3) // 4 * 4, dtype=torch.uint8))
        t.numpy()[:len(hv.pac)] = hv.pac
        pac_h[v] = t
        osz_h[v] = env.pinned(torch.from_numpy(hv.offset_size.reshape(-1).copy()))

    # every rank learns every volume's shape
    meta = torch.zeros((V, 4), dtype=torch.int64, device=dev)
    for v in own_vols:
        hv = host_volumes[v]
        meta[v] = torch.tensor([hv.num_reads, hv.num_bases, hv.start_read_id, pac_h[v].numel()], dtype=torch.int64)
    dist.all_reduce(meta, op=dist.ReduceOp.SUM)
    metas = [tuple(int(x) for x in m.tolist()) for m in meta.cpu()]
    max_pac = max(m[3] for m in metas); max_reads = max(m[0] for m in metas)
    reads_in_volume = [m[0] for m in metas]
    mirror = world - 1 - rank
    params = mecat_b200.pw_params(task=1)

    # Packed volumes in device memory: slot v of `pac_all` / `osz_all` holds volume v.  A rank fills the slots of its own
    # volumes from pinned host memory; the others arrive over NVLink in ONE NCCL all-gather per step (NCCL's ring algorithm
    # rotates the blocks from rank to rank), posted before the own indices are built and waited for after.  The first
    # version rotated the sets with a send/recv pair per ring step, double buffered behind the step's tiles: a rank could
    # not start step t+1 before its left neighbour had started step t, and the tiles per step are uneven (rank 3 of 8 has
    # none in steps 1-3 and two in steps 4-7), so every step cost the slowest rank's time: 4.5x on 8 GPUs
    # (profiles/r2_bench_ring_n8_stepwise.json).  Transport and compute are independent now.
    pac_all = torch.zeros((V, max_pac), dtype=torch.uint8, device=dev)
    osz_all = torch.zeros((V, max_reads * 2), dtype=torch.int32, device=dev)
    own_lo, own_hi = own_vols[0], own_vols[-1] + 1

    def dvolume(v):
        nr, nb, sid, _ = metas[v]
        osz = osz_all[v][:2 * nr].cpu().numpy().reshape(-1, 2)
        return ctx.volume_from_device(nr, nb, sid, osz, pac_all[v].data_ptr() if env.cuda else pac_all[v])

    served = sorted(set(sets[rank]) | set(sets[mirror]))
    items = [it for step in range(world) for it in ring_work(world, rank, step, V, reads_in_volume)]
    needed = sorted({s for s, _, _, _ in items} | {v for _, v, _, _ in items})
    io = {"h2d": 0, "nvlink": 0}
    phase = {"upload": 0.0, "index": 0.0, "gather_wait": 0.0, "volumes": 0.0, "tiles": 0.0}
    uploaded = [False]

    def clock(name, t0):
        t1 = time.perf_counter()
        phase[name] += (t1 - t0) * 1e3
        return t1

    def one_step(e2e):
        """Whole job once.  Returns the number of records this rank produced."""
        t0 = time.perf_counter()
        # own volumes to the device (inside the timed region only for e2e; resident otherwise)
        if e2e or not uploaded[0]:
            for v in own_vols:
                pac_all[v][:pac_h[v].numel()].copy_(pac_h[v], non_blocking=True)
                osz_all[v][:osz_h[v].numel()].copy_(osz_h[v], non_blocking=True)
                io["h2d"] += pac_h[v].numel() + osz_h[v].numel() * 4
            uploaded[0] = True
        handles = []
        if world > 1:
            env.sync()
            handles.append(dist.all_gather_into_tensor(pac_all.view(-1), pac_all[own_lo:own_hi].reshape(-1).clone(), async_op=True))
            handles.append(dist.all_gather_into_tensor(osz_all.view(-1), osz_all[own_lo:own_hi].reshape(-1).clone(), async_op=True))
            io["nvlink"] += (world - 1) * p * (max_pac + max_reads * 8)
        env.sync() if world == 1 else None
        dvols, idxs = {}, {}
        t0 = clock("upload", t0)
        if world == 1:
            for v in own_vols:
                dvols[v] = dvolume(v)
        else:
            # the own slots are not written by the gather (NCCL copies the send buffer into them, same bytes): safe to read
            for v in own_vols:
                dvols[v] = dvolume(v)
        t0 = clock("volumes", t0)
        for v in own_vols:
            if v in served:
                idxs[v] = ctx.index_build(dvols[v])
        t0 = clock("index", t0)
        for h in handles:
            h.wait()
        env.sync()
        t0 = clock("gather_wait", t0)
        for v in needed:
            if v not in dvols:
                dvols[v] = dvolume(v)
        t0 = clock("volumes", t0)
        for v in served:
            if v not in idxs:
                idxs[v] = ctx.index_build(dvols[v])
        t0 = clock("index", t0)
        produced = 0
        for s, v, rb, re in items:
            rec = ctx.pw_tile_range(idxs[s], dvols[s], dvols[v], params, rb, re)
            produced += len(rec)
        t0 = clock("tiles", t0)
        for i in idxs.values():
            ctx.release_index(i)
        for dv in dvols.values():
            ctx.release_volume(dv)
        return produced

    for i in range(args.warmup):
        n, dt = env.timed(1, one_step, False)
        if rank == 0:
            log("warmup %d: %d pairs in %.2f s" % (i, n, dt))
    ctx.reset_stats()
    for k in phase:
        phase[k] = 0.0
    sampler = ClockSampler(getattr(env, "local", 0)) if rank == 0 else None
    pairs, dt = env.timed(args.steps, one_step, False)
    stats = ctx.stats()
    phase_ms = {k: round(v / args.steps, 3) for k, v in phase.items()}
    busy = torch.tensor([sum(phase.values()) / args.steps], dtype=torch.float64, device=dev)
    busy_all = [torch.zeros_like(busy) for _ in range(world)]
    dist.all_gather(busy_all, busy)
    clocks = sampler.stop() if sampler else None
    esteps = max(1, min(args.steps, 2))
    ctx.reset_stats()
    io["h2d"] = io["nvlink"] = 0
    epairs, edt = env.timed(esteps, one_step, True)
    estats = ctx.stats()
    iot = torch.tensor([estats["h2d_bytes"] + io["h2d"], estats["d2h_bytes"], io["nvlink"]], dtype=torch.int64, device=dev)
    dist.all_reduce(iot, op=dist.ReduceOp.SUM)
    line = None
    if rank == 0:
        peaks = measured_peaks()
        tiles = V * (V + 1) // 2
        cfg = {"workload": "mecat2pw -j 1 all-vs-all, %d volumes x %d synthetic PacBio-CLR reads (15 kb mean, 15%% error, 15x "
                           "over a %d Mb genome) = %d reads, %d tiles (BASELINE configs[4] shape; equal volumes instead of the "
                           "splitter's 7 full volumes + a sliver)" % (V, READS, V * GENOME // 1000000, V * READS, tiles),
               "reads": V * READS, "genome": V * GENOME, "seed": SEED, "volumes": V, "tiles": tiles,
               "params": "-n 100 -a 2000 -k 4 -x 0", "l2": "inputs larger than L2 (no flush)",
               "parallelism": "1 gpu, all tiles" if world == 1 else
               "%d gpus: %d volume(s) per rank, packed volumes travel rank to rank in one NCCL all-gather (ring) per step, "
               "rank g serves the indices of ranks g and %d-g (mirror pairing), no other exchange" % (world, p, world - 1)}
        if args.reads:
            cfg["workload"] = "REDUCED debug workload: " + cfg["workload"]
        line = {
            "metric": METRIC, "value": pairs / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": epairs / edt, "unit": UNIT, "h2d_bytes_per_step": int(iot[0].item()) // esteps,
                    "d2h_bytes_per_step": int(iot[1].item()) // esteps, "nvlink_bytes_per_step": int(iot[2].item()) // esteps,
                    "ms_per_step": 1000.0 * edt / esteps, "steps": esteps},
            "gpu_launches": stats["gpu_launches"] * world,
            "roofline": roofline_for(stats, peaks, clocks=clocks),
            "cpu_baseline": {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                             "sample": "measured at N=1 only (bench.py --gpus 1)"},
            "pairs_per_step": pairs // args.steps,
            "phase_ms_per_step_rank0": phase_ms, "busy_ms_per_step_by_rank": [round(float(b.item()), 1) for b in busy_all],
            "kernel_ms_per_step_rank0": {k: round(v / args.steps, 3) for k, v in stats["kernel_ms"].items()},
        }
        print(json.dumps(line))
    ctx.close()
    dist.destroy_process_group()
    return line

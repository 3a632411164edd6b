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


def code_slices(world):
    """Equal slices of the 2^26 k-mer codes, aligned to 256 codes (the index kernels' granularity)."""
    n = 1 << 26
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


# ------------------------------------------------------------------------------------------ benchmark (strong scaling)
def run_bench_strong(args, METRIC, UNIT, workload_config, make_reads, tmp_root, ClockSampler, cpu_sample, roofline_for):
    """N GPUs share ONE tile (BASELINE configs[1]): every rank holds the packed volume, builds the
    slice [code_lo, code_hi) of the k-mer index, the slices are exchanged over NCCL (all-gather of the
    histogram, broadcast of every position slice), and rank r seeds / extends reads r*n/N..(r+1)*n/N."""
    import torch
    import torch.distributed as dist
    import mecat_b200

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def log(*a):
        print("[bench r%d]" % rank, *a, file=sys.stderr, flush=True)

    READS, GENOME, SEED = 100000, 100000000, 11
    if args.reads:
        READS = args.reads; GENOME = args.reads * 1000
    d = tmp_root()
    fa = os.path.join(d, "reads_%d_%d.fa" % (READS, SEED))
    wrk = os.path.join(d, "wrk_%d" % READS)
    if rank == 0:
        make_reads(fa, READS, GENOME, SEED)
        mecat_b200.split_dataset(fa, wrk)
    dist.barrier()
    vol = mecat_b200.HostVolume.load(os.path.join(wrk, "vol0"))
    pac = torch.empty(len(vol.pac), dtype=torch.uint8, pin_memory=True)
    pac.numpy()[:] = vol.pac
    osz = torch.from_numpy(vol.offset_size.reshape(-1).copy()).pin_memory()
    hv = mecat_b200.HostVolume(osz.numpy().reshape(-1, 2), pac.numpy(), vol.num_bases, 0)
    hv._keep = (pac, osz)
    params = mecat_b200.pw_params(task=1)
    ctx = mecat_b200.Context(local)
    lo, hi = code_slices(world)[rank]
    rb, re = read_slices(world, vol.num_reads)[rank]
    cuts = torch.tensor([c[0] for c in code_slices(world)] + [1 << 26], dtype=torch.int64, device=dev)
    resident = [None]
    # index exchange: "allgather" = one padded all-gather of the position slices (all NVLink ports busy at once) followed by
    # a device-side compaction; "broadcast" = one broadcast per slice (the first implementation, kept for comparison)
    exchange = os.environ.get("MECAT_STRONG_EXCHANGE", "allgather")
    phase = {"count": 0.0, "counts_allgather": 0.0, "finish": 0.0, "positions_exchange": 0.0, "tile": 0.0}

    def lap(name, t0):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        phase[name] += (t1 - t0) * 1e3
        return t1

    def one_step(e2e):
        t0 = time.perf_counter()
        if e2e or resident[0] is None:
            if resident[0] is not None:
                ctx.release_volume(resident[0])
            resident[0] = ctx.upload(hv)
        dvol = resident[0]
        idx = ctx.index_count_part(dvol, lo, hi)
        t0 = lap("count", t0)
        cptr, bptr, _, _ = ctx.index_device_arrays(idx)
        counts = device_view(cptr, 1 << 26, torch.int32, 4)
        if world > 1:
            parts = [counts[a:b] for a, b in code_slices(world)]
            if len({p.numel() for p in parts}) == 1:
                dist.all_gather_into_tensor(counts, parts[rank].clone())
            else:
                for q in range(world):
                    dist.broadcast(parts[q], src=q)
            torch.cuda.synchronize()
        t0 = lap("counts_allgather", t0)
        ctx.index_finish_part(dvol, idx, lo, hi)
        t0 = lap("finish", t0)
        if world > 1:
            _, bptr, pptr, nk = ctx.index_device_arrays(idx)
            begin = device_view(bptr, (1 << 26) + 1, torch.int32, 4)
            bounds = begin[cuts].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            pos = device_view(pptr, nk, torch.int32, 4)
            exchange_slices(dist, pos, bounds, rank, world, exchange)
            torch.cuda.synchronize()
        t0 = lap("positions_exchange", t0)
        rec = ctx.pw_tile_range(idx, dvol, dvol, params, rb, re)
        ctx.release_index(idx)
        lap("tile", t0)
        return len(rec)

    def timed(nsteps, e2e):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        for _ in range(nsteps):
            n += one_step(e2e)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        return int(c.item()), float(t.item())

    for i in range(args.warmup):
        n, dt = timed(1, False)
        if rank == 0:
            log("warmup %d: %d pairs in %.3f s" % (i, n, dt))
    ctx.reset_stats()
    for k in phase:
        phase[k] = 0.0
    sampler = ClockSampler(local) if rank == 0 else None
    pairs, dt = timed(args.steps, False)
    stats = ctx.stats()
    phase_ms = {k: round(v / args.steps, 3) for k, v in phase.items()}
    clocks = sampler.stop() if sampler else None
    esteps = max(1, min(args.steps, 3))
    ctx.reset_stats()
    epairs, edt = timed(esteps, True)
    estats = ctx.stats()
    io = torch.tensor([estats["h2d_bytes"], estats["d2h_bytes"]], dtype=torch.int64, device=dev)
    dist.all_reduce(io, op=dist.ReduceOp.SUM)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        except Exception:
            pass
        cfg = workload_config(1)
        cfg["parallelism"] = ("%d gpus sharing one tile: k-mer index built in %d code slices exchanged over NCCL "
                              "(all-gather), query reads split %d ways" % (world, world, world))
        if args.reads:
            cfg = dict(cfg, reads=READS, genome=GENOME, workload="REDUCED debug workload (%d reads)" % READS)
        line = {
            "metric": METRIC, "value": pairs / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": epairs / edt, "unit": UNIT, "h2d_bytes_per_step": int(io[0].item()) // esteps,
                    "d2h_bytes_per_step": int(io[1].item()) // esteps, "ms_per_step": 1000.0 * edt / esteps, "steps": esteps},
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


# ------------------------------------------------------------------------------------------ benchmark (block rotation)
def run_bench(args, METRIC, UNIT, workload_config, make_reads, tmp_root, ClockSampler, cpu_sample, roofline_for):
    import torch
    import torch.distributed as dist
    import mecat_b200

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def log(*a):
        print("[bench r%d]" % rank, *a, file=sys.stderr, flush=True)

    READS, GENOME, SEED = 100000, 100000000, 11
    if args.reads:
        READS = args.reads; GENOME = args.reads * 1000
    d = tmp_root()
    fa = os.path.join(d, "reads_w%d_r%d_%d_%d.fa" % (world, rank, READS, SEED))
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "gen_reads")
    if not (os.path.exists(fa) and os.path.getsize(fa) > READS * 2000):
        import subprocess
        # reads [rank*READS, (rank+1)*READS) of the N*READS-read data set over the N*GENOME genome
        subprocess.check_call([exe, fa + ".tmp", str(READS), str(GENOME * world), str(SEED), "15000", "1500", "0.15", "-",
                               str(rank * READS)])
        os.replace(fa + ".tmp", fa)
    wrk = os.path.join(d, "wrk_w%d_r%d_%d" % (world, rank, READS))
    names = mecat_b200.split_dataset(fa, wrk)
    assert len(names) == 1
    vol = mecat_b200.HostVolume.load(names[0])
    vol.start_read_id = rank * READS
    os.remove(fa)
    # pinned host copies (source of the H2D inside the e2e region)
    pac_h = torch.empty((len(vol.pac) + 3) // 4 * 4, dtype=torch.uint8, pin_memory=True)
    pac_h.zero_(); pac_h.numpy()[:len(vol.pac)] = vol.pac
    osz_h = torch.from_numpy(vol.offset_size.reshape(-1).copy()).pin_memory()

    # every rank learns every block's shape
    meta = torch.tensor([vol.num_reads, vol.num_bases, vol.start_read_id, pac_h.numel()], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [tuple(int(x) for x in m.tolist()) for m in metas]
    max_pac = max(m[3] for m in metas); max_reads = max(m[0] for m in metas)
    reads_in_block = [m[0] for m in metas]
    nxt, prv = ring_neighbours(world, rank)
    params = mecat_b200.pw_params(task=1)
    ctx = mecat_b200.Context(local)

    class Block:
        def __init__(self):
            self.pac = torch.empty(max_pac, dtype=torch.uint8, device=dev)
            self.osz = torch.empty(max_reads * 2, dtype=torch.int32, device=dev)
            self.v = -1

    bufs = [Block(), Block(), Block(), Block()]     # own block, two ring buffers, the mirror's block

    class Handle:
        def __init__(self, reqs, blk, v):
            self.reqs, self.blk, self.v = reqs, blk, v

        def wait(self):
            for r in self.reqs:
                r.wait()
            self.blk.v = self.v

    def exchange_with(dst, src, send_blk, recv_blk, recv_v):
        ops = [dist.P2POp(dist.isend, send_blk.pac, dst), dist.P2POp(dist.isend, send_blk.osz, dst),
               dist.P2POp(dist.irecv, recv_blk.pac, src), dist.P2POp(dist.irecv, recv_blk.osz, src)]
        return Handle(dist.batch_isend_irecv(ops), recv_blk, recv_v)

    def dvolume_of(blk):
        nr, nb, sid, _ = metas[blk.v]
        torch.cuda.synchronize()
        osz = blk.osz[:2 * nr].cpu().numpy().reshape(-1, 2)
        return ctx.volume_from_device(nr, nb, sid, osz, blk.pac.data_ptr())

    def one_step(e2e):
        """Whole job once.  Returns the number of records this rank produced."""
        own, ring_a, ring_b, mir = bufs
        # own block to the device (inside the timed region only for e2e; resident otherwise)
        if e2e or own.v != rank:
            own.pac[:pac_h.numel()].copy_(pac_h, non_blocking=True)
            own.osz[:osz_h.numel()].copy_(osz_h, non_blocking=True)
            own.v = rank
        mirror = world - 1 - rank
        dvols, idxs = {}, {}
        h = None
        if mirror != rank:
            h = exchange_with(mirror, mirror, own, mir, mirror)
        dvols[rank] = dvolume_of(own)
        idxs[rank] = ctx.index_build(dvols[rank])
        if h is not None:
            h.wait()
            dvols[mirror] = dvolume_of(mir)
            idxs[mirror] = ctx.index_build(dvols[mirror])
        produced = [0]

        def exchange(cur, sp):
            step_v = (cur.v - 1) % world
            return exchange_with(nxt, prv, cur, sp, step_v)

        def compute(step, blk):
            v = blk.v
            assert v == block_at(world, rank, step)
            items = tile_work(world, rank, step, reads_in_block)
            if not items:
                return
            dq = dvols[v] if v in dvols else dvolume_of(blk)
            for s, vv, rb, re in items:
                rec = ctx.pw_tile_range(idxs[s], dvols[s], dq, params, rb, re)
                produced[0] += len(rec)
            if v not in dvols:
                ctx.release_volume(dq)

        run_ring(world, rank, own, ring_a, ring_b, exchange, compute)
        for i in idxs.values():
            ctx.release_index(i)
        for dv in dvols.values():
            ctx.release_volume(dv)
        return produced[0]

    def timed(nsteps, e2e):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        for _ in range(nsteps):
            n += one_step(e2e)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        return int(c.item()), float(t.item())

    bufs[0].v = -1
    for i in range(args.warmup):
        n, dt = timed(1, False)
        if rank == 0:
            log("warmup %d: %d pairs in %.2f s" % (i, n, dt))
    ctx.reset_stats()
    sampler = ClockSampler(local) if rank == 0 else None
    pairs, dt = timed(args.steps, False)
    stats = ctx.stats()
    clocks = sampler.stop() if sampler else None
    esteps = max(1, min(args.steps, 2))
    ctx.reset_stats()
    epairs, edt = timed(esteps, True)
    estats = ctx.stats()
    h2d = torch.tensor([estats["h2d_bytes"] + pac_h.numel() * esteps + osz_h.numel() * 4 * esteps, estats["d2h_bytes"]],
                       dtype=torch.int64, device=dev)
    dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        except Exception:
            pass
        cfg = workload_config(world)
        if args.reads:
            cfg = dict(cfg, reads=world * READS, genome=world * GENOME, workload="REDUCED debug workload (%d reads per rank)" % READS)
        line = {
            "metric": METRIC, "value": pairs / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": epairs / edt, "unit": UNIT, "h2d_bytes_per_step": int(h2d[0].item()) // esteps,
                    "d2h_bytes_per_step": int(h2d[1].item()) // esteps, "ms_per_step": 1000.0 * edt / esteps, "steps": esteps},
            "gpu_launches": stats["gpu_launches"] * world,
            "roofline": roofline_for(stats, peaks, clocks=clocks),
            "cpu_baseline": {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                             "sample": "measured at N=1 only (bench.py --gpus 1)"},
            "pairs_per_step": pairs // args.steps,
            "kernel_ms_per_step_rank0": {k: round(v / args.steps, 3) for k, v in stats["kernel_ms"].items()},
            "nvlink_bytes_per_step_per_rank": int((world - 1 + (1 if world - 1 - rank != rank else 0)) * (max_pac + max_reads * 8)),
        }
        print(json.dumps(line))
    ctx.close()
    dist.destroy_process_group()

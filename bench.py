#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native mecat2pw hot path.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): overlapped read-pairs / second of `mecat2pw` (default job -j 1: block
k-mer seeding -> DDF candidate scoring -> O(nd) diff extension), i.e. M4 records produced per
second of wall time, whole job, on synthetic PacBio-CLR reads (15 kb, 15 % error).

A "step" is one pass of the hot path over the whole workload:
  N = 1   BASELINE configs[1]: all-vs-all on 100 000 x 15 kb reads (one 1.59 Gbase volume):
          k-mer index build + seeding/scoring of every read + extension of every candidate.
  N > 1   default (--mode strong): the SAME configs[1] tile shared by N GPUs: every rank builds one
          slice of the k-mer index, slices are exchanged over NCCL, query reads are split N ways.
          --mode ring: N such volumes (N x 100 000 reads, genome N x 100 Mb), all N(N+1)/2 tiles;
          packed query volumes rotate round the ring of ranks over NCCL/NVLink, rank g serves
          indices g and N-1-g (BASELINE configs[4] style; mecat_b200/multi.py).
`value`  = records / step time with the packed volume(s) already resident in HBM.
`e2e`    = the same through the host-buffer C-ABI call (mecat_b200_pw_overlaps): pinned host
           volume -> H2D -> index -> tile -> D2H records, all inside the timed region.
`--impl reference` times the UNMODIFIED reference binary (oracle/_ref/mecat2pw -j 1, all host
threads) on a bounded sample of the same read model (see sample_size(), cpu_sample()); its `config` names that sample
and `cpu_baseline.full_size` carries the committed full-size run of the same binary.
Inputs (>= 400 MB packed + 6 GB index) are far larger than the 126 MB L2: no flush needed.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READS_PER_VOLUME = 100000
GENOME_PER_VOLUME = 100000000
SEED = 11
METRIC = "overlapped read-pairs/sec (mecat2pw)"
UNIT = "pairs/s"
REFERENCE_BUDGET_S = 240.0    # wall-clock budget of the reference arm's repetitions (the sample is never shrunk to fit it)


def sample_size(cores):
    """Bounded CPU sample: same read model and 15x coverage as the workload.  The reference hands reads to its threads in
    chunks of 500 (CHUNK_SIZE, src/mecat2pw/pw_impl.h:15), so a sample needs >= 4 chunks per host thread or the threads
    starve: 2 000 reads per thread, at least 20 000, at most 40 000 reads (a 16-thread run of 32 000 reads takes ~1 min)."""
    reads = min(40000, max(20000, 2000 * cores))
    return reads, reads * 1000


def full_size_record():
    """The committed full-size run of the unmodified binary on the workload itself (tools/fullscale_parity.py)."""
    try:
        r = json.load(open(os.path.join(ROOT, "profiles", "r1_fullscale_parity_100k_j1.json")))
        return {"pairs": r["ref_records"], "seconds": round(r["ref_cli_seconds"], 2), "cores": r["cores"],
                "value": round(r["ref_records"] / r["ref_cli_seconds"], 1), "unit": UNIT,
                "what": "oracle/_ref/mecat2pw -j 1 -t %d on all %d reads, command-line wall clock, recorded once on a GPU box "
                        "(profiles/r1_fullscale_parity_100k_j1.json); not re-measured by this run" % (r["cores"], r["reads"])}
    except Exception:
        return None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def tmp_root():
    d = os.environ.get("MECAT_BENCH_TMP") or os.path.join(tempfile.gettempdir(), "mecat_b200_bench")
    os.makedirs(d, exist_ok=True)
    return d


def gen_exe():
    exe = os.path.join(ROOT, "mecat_b200", "bin", "gen_reads")
    if not os.path.exists(exe):
        from mecat_b200 import build
        build.build()
    return exe


def make_reads(path, n, genome, seed):
    if os.path.exists(path) and os.path.getsize(path) > n * 2000:
        return
    t = time.time()
    subprocess.check_call([gen_exe(), path + ".tmp", str(n), str(genome), str(seed)])
    os.replace(path + ".tmp", path)
    log("[bench] generated %d reads (genome %d, seed %d) in %.1f s" % (n, genome, seed, time.time() - t))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for i, nme in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU reference
def cpu_sample(threads=None, keep=None, tech=0):
    """Runs the unmodified reference binary (oracle/_ref/mecat2pw -j 1 -t <all cores>) on the
    bounded sample and returns (pairs, seconds, cores, description)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "mecat2pw")
    cores = threads or os.cpu_count() or 1
    SAMPLE_READS, SAMPLE_GENOME = sample_size(cores)
    if tech == 1:
        # the X-drop aligner costs the CPU ~6x more per candidate: 8 000 reads (one chunk of 500 per thread) keep the arm within minutes
        SAMPLE_READS, SAMPLE_GENOME = 8000, 8000000
    desc = ("SAMPLE, not the full workload: unmodified mecat2pw -j 1 -t %d, command-line wall clock, on %d synthetic CLR reads "
            "(15 kb, 15%% err, genome %d, seed %d): same read model and 15x coverage as the workload, %d chunks of 500 reads per "
            "host thread" % (cores, SAMPLE_READS, SAMPLE_GENOME, SEED, SAMPLE_READS // 500 // cores))
    if not os.path.exists(exe):
        return None, None, cores, "oracle/_ref/mecat2pw not built", "unavailable"
    d = tmp_root()
    fa = os.path.join(d, "sample_%d_%d.fa" % (SAMPLE_READS, SEED))
    make_reads(fa, SAMPLE_READS, SAMPLE_GENOME, SEED)
    wrk = tempfile.mkdtemp(prefix="refwrk_", dir=d)
    out = os.path.join(wrk, "out.m4")
    t = time.perf_counter()
    subprocess.check_call([exe, "-j", "1", "-d", fa, "-o", out, "-w", os.path.join(wrk, "w"), "-t", str(cores)] + (["-x", "1"] if tech == 1 else []),
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dt = time.perf_counter() - t
    with open(out, "rb") as f:
        pairs = sum(1 for _ in f)
    if keep:
        shutil.copy(out, keep)
    shutil.rmtree(wrk, ignore_errors=True)
    return pairs, dt, cores, desc, "reference"


def run_reference(args):
    """Reference arm: the unmodified binary on the bounded sample.  The SAMPLE is sized for the host (sample_size) and the
    number of REPETITIONS is what gets capped: runs are timed until --steps is reached or REFERENCE_BUDGET_S is spent
    (one run of the sample takes about a minute; a warm-up run only when a run is short).  `config` names the sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, pairs = [], 0
    cores = os.cpu_count() or 1
    t_all = time.perf_counter()
    pairs, dt, cores, desc, kind = cpu_sample(tech=args.tech)
    if pairs is None:
        print(json.dumps({"impl": "reference", "unavailable": desc}))
        return
    log("[bench] reference run 0: %d pairs in %.2f s" % (pairs, dt))
    warm = 0
    if args.warmup > 0 and dt < 20.0:
        warm = 1                      # a short run: treat the first one as warm-up (page cache, read generation)
    else:
        times.append(dt)
    while len(times) < args.steps and (time.perf_counter() - t_all) + (times[-1] if times else dt) < REFERENCE_BUDGET_S:
        pairs, dt, cores, desc, kind = cpu_sample(tech=args.tech)
        times.append(dt)
        log("[bench] reference run %d: %d pairs in %.2f s" % (len(times) + warm - 1, pairs, dt))
    total = sum(times)
    value = pairs * len(times) / total
    sreads, sgenome = sample_size(cores)
    if args.tech == 1:
        sreads, sgenome = 8000, 8000000
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "repetitions": {"timed": len(times), "warmup": warm, "requested_steps": args.steps, "requested_warmup": args.warmup,
                        "why": "each run is a whole mecat2pw job on the sample (~1 min); repetitions are capped at %d s of wall "
                               "clock, the sample is not shrunk" % int(REFERENCE_BUDGET_S)},
        "config": {"workload": "SAMPLE of the GPU arm's workload: mecat2pw -j 1 all-vs-all on %d synthetic PacBio-CLR reads (15 kb mean, "
                               "15%% error, 15x) -- the GPU arm runs all %d reads; the CPU's full-size rate is in cpu_baseline.full_size"
                               % (sreads, READS_PER_VOLUME),
                   "reads": sreads, "genome": sgenome, "seed": SEED, "params": "-n 100 -a 2000 -k 4 -x 0" if args.tech == 0 else "-n 100 -a 500 -k 2 -x 1",
                   "parallelism": "reference CPU binary, %d host threads (no GPU)" % cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc, "full_size": full_size_record()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n):
    return {"workload": "mecat2pw -j 1 all-vs-all, %d x %d synthetic PacBio-CLR reads (15 kb mean, 15%% error, 15x), "
                        "%d volume(s), %d tile(s)" % (n, READS_PER_VOLUME, n, n * (n + 1) // 2),
            "reads": n * READS_PER_VOLUME, "genome": n * GENOME_PER_VOLUME, "seed": SEED,
            "params": "-n 100 -a 2000 -k 4 -x 0", "l2": "inputs larger than L2 (no flush)",
            "parallelism": "1 gpu" if n == 1 else "%d gpus: query-volume ring over NCCL, mirror-paired indices" % n}


# ------------------------------------------------------------------------------------------ ours
def pinned_volume(vol):
    """Copy the packed volume into pinned host memory (torch) so the H2D inside e2e is a DMA."""
    import numpy as np
    import torch
    import mecat_b200
    pac = torch.empty(len(vol.pac), dtype=torch.uint8, pin_memory=True)
    pac.numpy()[:] = vol.pac
    osz = torch.empty(vol.offset_size.size, dtype=torch.int32, pin_memory=True)
    osz.numpy()[:] = vol.offset_size.reshape(-1)
    hv = mecat_b200.HostVolume(osz.numpy().reshape(-1, 2), pac.numpy(), vol.num_bases, vol.start_read_id)
    hv._keep = (pac, osz)
    assert hv.pac.ctypes.data == pac.data_ptr()
    return hv


def ncu_counters():
    """Per-unit counters taken from the committed ncu captures of this command (profiles/ncu_counters.json): DRAM bytes per
    step and warp-instructions per extension block / per index hit.  They cannot be measured outside a profiler, so the
    bench multiplies them with the unit counts and event times it measures live."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_counters.json")))
    except Exception:
        return {}


def roofline_for(stats, peaks, nsteps=None, world=1, clocks=None):
    """Roofline entries for the timed steps.  `world` > 1: the caller's statistics are one rank's
    (strong mode); the index kernels then work on 1/world of the k-mer codes and every algorithmic figure is per rank.

    * the entry itself: the dominant kernel against the HBM roof (the contract's form);
    * "issue": the same kernel against the instruction-issue roof it actually runs into (4 warp schedulers per SM, one
      warp-instruction per cycle each), plus furthest-point cells per second;
    * "hbm_kernel": the largest kernel that IS bound by memory traffic (the hit-streaming `seed` kernel)."""
    km = stats["kernel_ms"]
    launches = stats["kernel_launches"]
    name = max(km, key=lambda k: km[k])
    B = stats["index_bases"]        # summed over steps
    K = stats["index_kmers"]
    H = stats["num_hits"]
    C = stats["num_candidates"]
    ncodes = 1 << 26
    steps = nsteps or max(1, launches["index_scan"] // 3 or launches["index_count"])     # index builds in the timed region
    w = float(max(1, world))
    algs = {
        # bytes the algorithm must move, summed over the timed steps; a rank of a strong-scaling run
        # reads every packed base but histograms / fills / sorts only its slice of the codes
        "index_count": B / 4 + (4.0 * K + 4.0 * ncodes * steps) / w,      # packed bases in, one counter update per k-mer
        "index_fill": B / 4 + (4.0 * K + 8.0 * ncodes * steps) / w,       # packed bases in, begin[] in, one position out per kept k-mer
        "index_sort": (8.0 * K + 4.0 * ncodes * steps) / w,               # positions in and out, begin[] in
        "index_scan": 12.0 * ncodes * steps / w,
        "seed": 3 * 4.0 * H + 0.0,                                        # three streaming passes over the hit positions
        "extend": C * (2 * 15000 / 4 + 52 + 32),                          # two packed reads in, one record out per candidate
    }
    cnt = ncu_counters()

    def entry(kname):
        alg = algs.get(kname, 0.0)
        ms = km[kname]
        achieved = alg / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        peak = peaks.get("hbm_gbs", 6650.0)
        # DRAM traffic from the committed `ncu --set full` capture of this command at N = 1 (bytes per step of the whole
        # workload), scaled to one launch like `achieved`; a rank of an N-GPU run moves 1/N of it
        traffic = None
        per_step = cnt.get(kname, {}).get("dram_bytes_per_step")
        if per_step and launches[kname]:
            traffic = per_step * steps / w / launches[kname]
        return {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "ms_per_launch": ms / max(1, launches[kname]), "launches": launches[kname],
                "algorithmic_bytes_per_launch": alg / max(1, launches[kname]),
                "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"}

    roof = entry(name)
    peak = roof["peak"]
    every = {}
    for k, a in algs.items():
        if km.get(k, 0) > 0:
            g = a / (km[k] * 1e-3) / 1e9
            every[k] = {"ms_per_step": round(km[k] / steps, 3), "algorithmic_gb_per_step": round(a / steps / 1e9, 3),
                        "gbps": round(g, 1), "frac_of_hbm_peak": round(g / peak, 4) if peak else None}
    roof["all_kernels"] = every
    roof["kernel_ms_share"] = {k: round(v / max(1e-9, sum(km.values())), 4) for k, v in km.items() if v > 0}
    # the issue roof of the extension kernel
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    sms = 148
    blocks = stats.get("num_extend_blocks", 0)
    cells = stats.get("num_extend_cells", 0)
    wipb = cnt.get("extend", {}).get("warp_inst_per_block")
    if km.get("extend", 0) > 0 and blocks:
        sec = km["extend"] * 1e-3
        issue_peak = 4.0 * sms * sm_mhz * 1e6
        iss = {"kernel": "extend", "bound": "issue", "unit": "warp-instructions/s", "peak": issue_peak,
               "peak_source": "4 schedulers x %d SMs x %.0f MHz (SM clock sampled during the timed region)" % (sms, sm_mhz),
               "blocks_per_s": blocks / sec, "cells_per_s": cells / sec if cells else None,
               "warp_inst_per_block": wipb, "warp_inst_source": cnt.get("extend", {}).get("source")}
        if wipb:
            iss["achieved"] = wipb * blocks / sec
            iss["frac"] = iss["achieved"] / issue_peak
        roof["issue"] = iss
    if name != "seed" and km.get("seed", 0) > 0:
        roof["hbm_kernel"] = entry("seed")
    return roof


def run_ours(args):
    import numpy as np
    import torch
    import mecat_b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 or args.mode == "ring":
        from mecat_b200 import multi
        if world == 1:      # a single process still goes through torch.distributed (one-rank NCCL group)
            os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1"); os.environ.setdefault("LOCAL_RANK", "0")
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29581")
        fn = multi.run_bench if args.mode == "ring" else multi.run_bench_strong
        return fn(args, METRIC, UNIT, workload_config, make_reads, tmp_root, ClockSampler, cpu_sample, roofline_for)
    torch.cuda.set_device(local)
    d = tmp_root()
    nreads = args.reads or (20000 if args.tech == 1 else READS_PER_VOLUME)
    genome = int(GENOME_PER_VOLUME * (nreads / READS_PER_VOLUME))
    fa = os.path.join(d, "reads_%d_%d.fa" % (nreads, SEED))
    make_reads(fa, nreads, genome, SEED)
    wrk = os.path.join(d, "wrk_%d" % nreads)
    t = time.time()
    names = mecat_b200.split_dataset(fa, wrk)
    assert len(names) == 1, "workload must fit one volume"
    vol = mecat_b200.HostVolume.load(names[0])
    hv = pinned_volume(vol)
    log("[bench] split + load: %.1f s; %d reads, %d bases" % (time.time() - t, vol.num_reads, vol.num_bases))
    params = mecat_b200.pw_params(task=1) if args.tech == 0 else mecat_b200.pw_params(task=1, min_align_size=500, min_kmer_match=2, tech=1)
    ctx = mecat_b200.Context(local)
    dvol = ctx.upload(hv)

    digests = set()

    def digest(rec):
        # order-sensitive fingerprint of the records of one step (determinism check across steps)
        w = np.frombuffer(rec.tobytes(), dtype=np.uint64)
        return (len(rec), int(np.bitwise_xor.reduce(w * np.arange(1, len(w) + 1, dtype=np.uint64))) if len(w) else 0)

    def step_resident(check=False):
        idx = ctx.index_build(dvol)
        rec = ctx.pw_tile(idx, dvol, dvol, params)
        ctx.release_index(idx)
        if check:
            digests.add(digest(rec))
        return len(rec)

    def step_e2e(check=False):
        rec = ctx.pw_overlaps(hv, hv, params)
        if check:
            digests.add(digest(rec))
        return len(rec)

    for i in range(args.warmup):
        t = time.perf_counter()
        n = step_resident(check=True)
        log("[bench] warmup %d: %d pairs in %.2f s" % (i, n, time.perf_counter() - t))
    ctx.reset_stats()
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pairs = 0
    for i in range(args.steps):
        pairs += step_resident()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stats = ctx.stats()
    clocks = sampler.stop()
    log("[bench] resident: %d pairs in %.2f s; kernel ms %s" % (pairs, dt, json.dumps(stats["kernel_ms"])))
    # end to end through the host-buffer C-ABI call
    step_e2e(check=True)
    ctx.reset_stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    epairs = 0
    esteps = max(1, min(args.steps, 3))
    for i in range(esteps):
        epairs += step_e2e()
    torch.cuda.synchronize()
    edt = time.perf_counter() - t0
    estats = ctx.stats()
    ctx.release_volume(dvol)
    ctx.close()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roof = roofline_for(stats, peaks, args.steps, clocks=clocks)
    cb = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "skipped (--no-cpu)"}
    if not args.no_cpu:
        cp, cdt, cores, desc, kind = cpu_sample(tech=args.tech)
        if cp is not None:
            cb = {"value": cp / cdt, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc, "seconds": cdt, "pairs": cp,
                  "full_size": full_size_record()}
        else:
            cb = {"value": None, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
    line = {
        "metric": METRIC, "value": pairs / dt, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": workload_config(1) if not args.reads and args.tech == 0 else
        dict(workload_config(1), reads=nreads, genome=genome, workload=("REDUCED debug workload (%d reads)" % nreads) if args.tech == 0 else
             "mecat2pw -j 1 -x 1 (nanopore parameter set: X-drop aligner, -a 500 -k 2, min_kmer_dist 400) all-vs-all on %d synthetic reads "
             "(15 kb mean, 15%% error, 15x); not the headline configuration" % nreads, params="-n 100 -a 500 -k 2 -x 1"),
        "clocks": clocks,
        "e2e": {"value": epairs / edt, "unit": UNIT, "h2d_bytes_per_step": estats["h2d_bytes"] // esteps,
                "d2h_bytes_per_step": estats["d2h_bytes"] // esteps, "ms_per_step": 1000.0 * edt / esteps, "steps": esteps},
        "gpu_launches": stats["gpu_launches"],
        "roofline": roof, "cpu_baseline": cb,
        "pairs_per_step": pairs // args.steps,
        "deterministic": len(digests) == 1,   # warm-up steps and the e2e call produced identical record arrays
        "kernel_ms_per_step": {k: round(v / args.steps, 3) for k, v in stats["kernel_ms"].items()},
        "host_ms_per_step": round(stats["host_ms"] / args.steps, 3), "d2h_ms_per_step": round(stats["d2h_ms"] / args.steps, 3),
        "wall_ms_per_step": {k: round(stats[k] / args.steps, 3) for k in ("wall_index_ms", "wall_seed_ms", "wall_extend_ms", "total_ms")},
        "hits_per_step": stats["num_hits"] // args.steps, "candidates_per_step": stats["num_candidates"] // args.steps,
        "extend_blocks_per_step": stats["num_extend_blocks"] // args.steps,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ the other two programs
def _cli(exe, argv, env=None, cwd=None):
    """Run a command-line driver; returns (wall seconds, stderr text)."""
    t = time.perf_counter()
    p = subprocess.run([exe] + argv, capture_output=True, text=True, env=dict(os.environ, **(env or {})), cwd=cwd)
    dt = time.perf_counter() - t
    if p.returncode != 0:
        raise SystemExit("%s failed: %s" % (exe, p.stderr[-1500:]))
    return dt, p.stderr


def run_program(args):
    """--workload ref | cns: BASELINE configs[2] / configs[3] through the command-line drivers (the product path of those
    programs).  `value` = reads per second of the device phase the driver itself times (mapping / consensus of all reads,
    inputs already packed), `e2e` = reads per second of the whole command line (process start to exit: FASTA in, text out).
    --impl reference: the unmodified binary, all host threads, on a bounded sample that its `config` names."""
    import re
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d = tmp_root()
    cores = os.cpu_count() or 1
    nreads, genome = READS_PER_VOLUME, GENOME_PER_VOLUME
    fa, gfa = os.path.join(d, "reads_%d_%d.fa" % (nreads, SEED)), os.path.join(d, "genome_%d_%d.fa" % (genome, SEED))
    if not (os.path.exists(gfa) and os.path.exists(fa) and os.path.getsize(fa) > nreads * 2000):
        subprocess.check_call([gen_exe(), fa + ".tmp", str(nreads), str(genome), str(SEED), "15000", "1500", "0.15", gfa])
        os.replace(fa + ".tmp", fa)
    bindir = os.path.join(ROOT, "mecat_b200", "bin")
    refdir = os.path.join(ROOT, "oracle", "_ref")
    ref = args.impl == "reference"
    if args.workload == "asm":
        # the corrected-read overlapper of mecat2canu: one block file of corrected-read like reads
        # against itself; the reference arm runs the same workload, not a sample (it takes ~20 s on 16 cores)
        metric, unit = "reads overlapped/sec (mecat2asmpw, one block file against itself)", "reads/s"
        nreads, genome = 20000, 2500000
        blocks = os.path.join(d, "asm_blocks_%d_%d" % (nreads, SEED))
        fa = os.path.join(blocks, "000001.fasta")
        if not (os.path.exists(fa) and os.path.exists(os.path.join(blocks, "ovlprep"))):
            os.makedirs(blocks, exist_ok=True)
            subprocess.check_call([gen_exe(), fa + ".tmp", str(nreads), str(genome), str(SEED), "4000", "800", "0.015", "-"])
            os.replace(fa + ".tmp", fa)
            with open(os.path.join(blocks, "ovlprep"), "w") as f:
                f.write("-allreads -allbases -b 1 -e %d\n" % nreads)
        for f in os.listdir(blocks):
            if f.endswith(".r"):
                os.remove(os.path.join(blocks, f))
        exe = os.path.join(refdir if ref else bindir, "mecat2asmpw")
        argv, units = ["-P" + blocks, "-T%d" % cores, "-S1", "-E1"], nreads
        workload = "mecat2asmpw -S1 -E1: %d x 4 kb synthetic corrected reads (1.5 %% error, 32x of a %.1f Mb genome), every read against the index of the file" % (nreads, genome / 1e6)
    elif args.workload == "ref":
        metric, unit = "reads mapped/sec (mecat2ref -m 1)", "reads/s"
        sample = min(nreads, max(8000, 500 * cores))
        if ref:
            sfa = os.path.join(d, "ref_sample_%d.fa" % sample)
            if not os.path.exists(sfa):
                with open(fa) as f, open(sfa, "w") as g:
                    k = 0
                    for line in f:
                        if line.startswith(">"):
                            k += 1
                            if k > sample:
                                break
                        g.write(line)
            exe, argv, units = os.path.join(refdir, "mecat2ref"), ["-d", sfa, "-r", gfa, "-o", os.path.join(d, "ref_ref.m4"), "-w", os.path.join(d, "wr"), "-t", str(cores), "-m", "1"], sample
        else:
            exe, argv, units = os.path.join(bindir, "mecat2ref"), ["-d", fa, "-r", gfa, "-o", os.path.join(d, "ref_gpu.m4"), "-w", os.path.join(d, "wg"), "-t", str(cores), "-m", "1"], nreads
        workload = "mecat2ref -m 1: %d x 15 kb synthetic CLR reads vs their %d Mb genome (BASELINE configs[2])" % (nreads, genome // 1000000)
    else:
        metric, unit = "reads corrected/sec (mecat2cns -i 0)", "reads/s"
        can = os.path.join(d, "cand_%d_%d.can" % (nreads, SEED))
        if not os.path.exists(can):       # candidates of the same reads: mecat2pw -j 0 on the GPU (set-up, not timed)
            subprocess.check_call([os.path.join(bindir, "mecat2pw"), "-j", "0", "-d", fa, "-o", can + ".tmp", "-w", os.path.join(d, "wcan")],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            os.replace(can + ".tmp", can)
        sample = min(nreads, max(4000, 400 * cores))
        if ref:
            scan = os.path.join(d, "cand_sample_%d.can" % sample)
            if not os.path.exists(scan):  # both orientations of a pair become templates: keep pairs with a template below `sample`
                with open(can) as f, open(scan, "w") as g:
                    for line in f:
                        a = line.split("\t", 2)
                        if int(a[0]) < sample or int(a[1]) < sample:
                            g.write(line)
            exe, argv, units = os.path.join(refdir, "mecat2cns"), ["-i", "0", "-t", str(cores), "-p", str(sample), scan, fa, os.path.join(d, "cns_ref.fa")], sample
        else:
            exe, argv, units = os.path.join(bindir, "mecat2cns"), ["-i", "0", "-t", str(cores), can, fa, os.path.join(d, "cns_gpu.fa")], nreads
        workload = "mecat2cns -i 0 on the candidates of mecat2pw -j 0 for %d x 15 kb synthetic CLR reads (BASELINE configs[3])" % nreads
    if not os.path.exists(exe):
        print(json.dumps({"impl": args.impl, "unavailable": "%s is not built" % exe}))
        return
    walls, phases, log = [], [], ""
    reps = args.warmup + args.steps
    t_all = time.perf_counter()
    for i in range(reps):
        if ref and walls and (time.perf_counter() - t_all) + walls[-1] > REFERENCE_BUDGET_S:
            break
        dt, log = _cli(exe, argv, env={"MECAT_B200_STATS": "1"}, cwd=d)
        ph = None
        if not ref:
            if args.workload == "asm":
                m = re.search(r"index ([0-9.]+) s, mapping ([0-9.]+) s", log)
                ph = float(m.group(1)) + float(m.group(2)) if m else None
            elif args.workload == "ref":
                m = re.search(r"mapping ([0-9.]+) s", log)
                ph = float(m.group(1)) if m else None
            else:
                ph = sum(float(x) for x in re.findall(r"processing reads .* takes ([0-9.]+) secs", log)) or None
        if i >= args.warmup or ref:
            walls.append(dt); phases.append(ph)
        print("[bench] %s run %d: %.2f s wall%s" % (args.workload, i, dt, "" if ph is None else ", device phase %.2f s" % ph), file=sys.stderr, flush=True)
    n = len(walls)
    wall = sum(walls) / n
    e2e_value = units / wall
    phase = (sum(phases) / n) if phases and all(x is not None for x in phases) else None
    value = units / phase if phase else e2e_value
    kernel = None
    m = re.search(r"\[kernel ms\](.*)", log)
    if m:
        kernel = {k: float(v) for k, v in re.findall(r"(\w+)=([0-9.]+)\(", m.group(1))}
    in_bytes = os.path.getsize(fa) + (0 if args.workload == "asm" else os.path.getsize(gfa) if args.workload == "ref" else os.path.getsize(os.path.join(d, "cand_%d_%d.can" % (nreads, SEED))))
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * (phase if phase else wall), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": ("SAMPLE (%d reads) of: " % units if ref and args.workload != "asm" else "") + workload, "reads": units, "genome": genome, "seed": SEED,
                   "parallelism": ("reference CPU binary, %d host threads (no GPU)" % cores) if ref else "1 gpu, command-line driver"},
        "e2e": {"value": e2e_value, "unit": unit, "ms_per_step": 1000.0 * wall, "steps": n,
                "h2d_bytes_per_step": 0 if ref else None, "d2h_bytes_per_step": 0 if ref else None,
                "what": "whole command line, process start to exit (%d bytes of input files read, text output written)" % in_bytes},
        "runs": {"wall_s": [round(x, 3) for x in walls], "device_phase_s": [None if x is None else round(x, 3) for x in phases],
                 "median_wall_s": round(statistics.median(walls), 3),
                 "note": "value / e2e are means over the timed runs; process start-up (CUDA context) makes single command-line runs on a shared box vary by seconds"},
        "gpu_launches": 0 if ref else None, "kernel_ms_last_run": kernel,
        "cpu_baseline": {"value": e2e_value if ref else None, "unit": unit, "cores": cores, "kind": "reference",
                         "sample": (("unmodified binary as the reference builds it (no CFLAGS), the whole workload" if args.workload == "asm" else "unmodified binary on the first %d reads / templates" % units)) if ref else "run bench.py --workload %s --impl reference" % args.workload},
    }
    if ref:
        line["impl"] = "reference"
        line["repetitions"] = {"timed": n, "requested_steps": args.steps}
        if args.workload == "ref":
            # the reference's own timers (config.txt in its working directory): the genome index is a fixed cost that a
            # sample of the reads does not shrink, so the full-size rate is estimated from them
            try:
                txt = open(os.path.join(d, "config.txt")).read()
                t_idx = float(re.search(r"Building Reference Index Time: ([0-9.]+)", txt).group(1))
                t_map = float(re.search(r"Mapping Time: ([0-9.]+)", txt).group(1))
                t_tot = float(re.search(r"total Time : ([0-9.]+)", txt).group(1))
                full = t_idx + (t_tot - t_idx) * nreads / units
                line["cpu_baseline"]["reference_timers_s"] = {"index": t_idx, "mapping": t_map, "total": t_tot}
                line["cpu_baseline"]["full_size_estimate"] = {"value": nreads / full, "unit": unit,
                                                              "how": "index time + (total - index) x %d / %d reads" % (nreads, units)}
            except Exception:
                pass
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=0, help="debug only: reduced workload (result is not the headline)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--tech", type=int, default=0, choices=[0, 1], help="-x of mecat2pw: 1 = the nanopore parameter set (workload pw only; 20 000 reads by default)")
    ap.add_argument("--workload", default="pw", choices=["pw", "ref", "cns", "asm"],
                    help="pw (default): the headline, mecat2pw -j 1 (BASELINE configs[1]).  ref / cns: mecat2ref (configs[2]) / "
                         "mecat2cns (configs[3]) through their command-line drivers")
    ap.add_argument("--mode", default="strong", choices=["strong", "ring"],
                    help="strong (default): N GPUs share the configs[1] tile. ring: the BASELINE configs[4] job -- --volumes "
                         "volumes (default 8 x 125 000 reads = 1 M reads, 36 tiles) at any N that divides it, volume sets "
                         "rotate over NCCL; also runs at N = 1 (the baseline of its 1 -> 8 scaling)")
    ap.add_argument("--volumes", type=int, default=0, help="--mode ring: number of volumes (default 8)")
    ap.add_argument("--ring-reads", type=int, default=0, help="--mode ring: reads per volume (default 125 000)")
    args = ap.parse_args()
    # Only the JSON line may reach stdout: libraries (NCCL prints its version there) go to stderr.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.workload != "pw":
        run_program(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()

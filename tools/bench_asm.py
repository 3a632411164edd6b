"""mecat2asmpw on one block file of corrected-read like reads: the CUDA path (C ABI and command line) next to the
unmodified binary on the box's host cores.

  python tools/bench_asm.py [--reads 20000] [--genome 2500000] [--mean 4000] [--steps 3] [--no-ref] > gpurun_out/bench_asm.json

Reads: tools/gen_reads.cpp at 1.5 % error (corrected reads), one block file, every read against the index of the file.
The reference is timed twice: as its own makefile builds it (no CFLAGS) and the same source with -O2, both with all host
threads (the program sleeps 2 s per batch in its thread start-up, mecat2asmpw.c:988-1000; that is part of its time).
Lines that differ between the two outputs are counted: where the binary reads block memory no seed of the strand wrote,
its result depends on which reads its thread mapped before (oracle/oracle_asmpw.cpp)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=20000)
    ap.add_argument("--genome", type=int, default=2500000)
    ap.add_argument("--mean", type=int, default=4000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    import mecat_b200
    import util
    tmp = tempfile.mkdtemp(prefix="bench_asm_")
    wrk = os.path.join(tmp, "blocks")
    os.makedirs(wrk)
    fa = os.path.join(wrk, "000001.fasta")
    util.gen_reads(fa, a.reads, a.genome, 11, a.mean, a.mean // 5, err=0.015)
    with open(os.path.join(wrk, "ovlprep"), "w") as f:
        f.write("-allreads -allbases -b 1 -e %d\n" % a.reads)
    seqs = [s for _, s in util.read_fasta_raw(fa)]
    letters = sum(len(s) for s in seqs)
    out = {"workload": "mecat2asmpw, one block file against itself", "reads": a.reads, "letters": letters, "genome": a.genome, "error_rate": 0.015}

    ctx = mecat_b200.Context(0)
    reads = mecat_b200.AsmReads(seqs, 1)
    times = []
    for step in range(a.steps + 1):
        ctx.reset_stats()
        t0 = time.time()
        idx = ctx.asm_index_build(reads)
        t1 = time.time()
        recs = ctx.asm_overlaps(idx, reads, 0, 100)
        t2 = time.time()
        ctx.asm_index_release(idx)
        if step or a.steps == 0:
            times.append((t1 - t0, t2 - t1))
        st = ctx.stats()
    ours = sorted(mecat_b200.asm_lines(recs))
    best = min(times, key=lambda x: x[0] + x[1])
    out["ours"] = {"index_s": best[0], "overlaps_s": best[1], "reads_per_s": a.reads / (best[0] + best[1]), "records": len(ours),
                   "kernel_ms": {k: v for k, v in st["kernel_ms"].items() if k.startswith("asm")},
                   "kernel_launches": {k: v for k, v in st["kernel_launches"].items() if k.startswith("asm")},
                   "hits": st["num_hits"], "candidates": st["num_candidates"], "steps": a.steps, "all_steps_s": times}
    ctx.close()
    t0 = time.time()
    p = subprocess.run([os.path.join(ROOT, "mecat_b200", "bin", "mecat2asmpw"), "-P" + wrk, "-T1", "-S1", "-E1"], capture_output=True, text=True, check=True)
    out["ours"]["command_line_s"] = time.time() - t0
    out["ours"]["command_line_phases"] = p.stderr.strip().splitlines()[-1]
    cli = sorted(open(os.path.join(wrk, "1_0.r")).read().splitlines())
    out["ours"]["command_line_equals_abi"] = cli == ours
    os.remove(os.path.join(wrk, "1_0.r"))

    if not a.no_ref:
        cores = os.cpu_count()
        for tag, exe in (("reference", "mecat2asmpw"), ("reference_O2", "O2_mecat2asmpw")):
            path = os.path.join(ROOT, "oracle", "_ref", exe)
            if not os.path.exists(path):
                continue
            for f in os.listdir(wrk):
                if f.endswith(".r"):
                    os.remove(os.path.join(wrk, f))
            t0 = time.time()
            subprocess.check_call([path, "-P" + wrk, "-T%d" % cores, "-S1", "-E1"])
            dt = time.time() - t0
            lines = []
            for f in os.listdir(wrk):
                if f.endswith(".r"):
                    lines += open(os.path.join(wrk, f)).read().splitlines()
            lines.sort()
            so, sr = set(ours), set(lines)
            out[tag] = {"seconds": dt, "reads_per_s": a.reads / dt, "cores": cores, "records": len(lines),
                        "lines_only_ours": len(so - sr), "lines_only_reference": len(sr - so), "identical": lines == ours}
        if "reference" in out:
            out["speedup_command_line"] = out["reference"]["seconds"] / out["ours"]["command_line_s"]
            out["speedup_abi"] = out["reference"]["seconds"] / (best[0] + best[1])
    print(json.dumps(out))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Parity + timing of `mecat2ref` at scale (BASELINE configs[2] shape: synthetic 15 kb CLR reads against their own
genome): the GPU driver on all reads, the unmodified reference binary (all host cores) on a bounded sample of the same
reads against the same genome; the sample's sorted records must be identical.  Writes gpurun_out/bench_ref_<reads>.json.
Test/bench tooling (executes oracle/_ref as checker and CPU baseline)."""
import argparse
import hashlib
import json
import os
import subprocess
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def m4_records(path, max_read=None):
    lines = open(path).read().splitlines()
    if max_read is not None:
        lines = [l for l in lines if int(l.split("\t", 1)[0]) < max_read]
    lines.sort()
    return len(lines), hashlib.sha256("\n".join(lines).encode()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100000)
    ap.add_argument("--coverage", type=int, default=15)
    ap.add_argument("--sample", type=int, default=4000, help="reads the reference binary maps (the first ones of the file)")
    ap.add_argument("--format", type=int, default=1)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--tmp", default="/tmp/mecat_bench_ref")
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--forward", action="store_true", help="m4 format: forward-only extensions (MECAT_B200_REF_EXTEND=forward)")
    ap.add_argument("--driver", default=os.path.join(ROOT, "mecat_b200", "bin", "mecat2ref"), help="mecat2ref executable under test")
    a = ap.parse_args()
    os.makedirs(a.tmp, exist_ok=True)
    fa, genome_fa = os.path.join(a.tmp, "reads.fa"), os.path.join(a.tmp, "genome.fa")
    genome = a.reads * 15000 // a.coverage
    subprocess.check_call([os.path.join(ROOT, "mecat_b200", "bin", "gen_reads"), fa, str(a.reads), str(genome), "11", "15000", "1500", "0.15", genome_fa])
    res = {"reads": a.reads, "genome": genome, "coverage": a.coverage, "cores": os.cpu_count(), "format": a.format, "gpus": a.gpus, "forward_only": a.forward}
    gout = os.path.join(a.tmp, "gpu.out")
    t = time.time()
    p = subprocess.run([a.driver, "-d", fa, "-r", genome_fa, "-o", gout, "-w", os.path.join(a.tmp, "wg"),
                        "-m", str(a.format)], capture_output=True, text=True, env=dict(os.environ, MECAT_GPUS=str(a.gpus), MECAT_B200_STATS="1", **({"MECAT_B200_REF_EXTEND": "forward"} if a.forward else {})))
    res["gpu_cli_seconds"] = time.time() - t
    res["gpu_log"] = p.stderr.splitlines()[-4:]
    assert p.returncode == 0, p.stderr[-2000:]
    res["gpu_reads_per_second_cli"] = a.reads / res["gpu_cli_seconds"]
    if a.format == 1:
        n, sha = m4_records(gout)
        res.update(gpu_records=n, gpu_sha256=sha)
    if not a.skip_ref and a.format == 1:
        sample = min(a.sample, a.reads)
        sfa = os.path.join(a.tmp, "sample.fa")
        with open(fa) as f, open(sfa, "w") as g:
            k = 0
            for line in f:
                if line.startswith(">"):
                    k += 1
                    if k > sample:
                        break
                g.write(line)
        rout = os.path.join(a.tmp, "ref.out")
        t = time.time()
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "mecat2ref"), "-d", sfa, "-r", genome_fa, "-o", rout, "-w", os.path.join(a.tmp, "wr"),
                               "-t", str(os.cpu_count()), "-m", "1"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=a.tmp)
        res["ref_cli_seconds"] = time.time() - t
        res["ref_sample_reads"] = sample
        res["ref_timing_lines"] = open(os.path.join(a.tmp, "config.txt")).read().splitlines()[-3:]
        n2, sha2 = m4_records(rout)
        n1, sha1 = m4_records(gout, max_read=sample)
        res.update(ref_records=n2, gpu_records_in_sample=n1, identical_on_sample=(sha1 == sha2 and n1 == n2))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_ref_%d.json" % a.reads), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

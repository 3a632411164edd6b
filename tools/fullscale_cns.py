#!/usr/bin/env python
"""Parity + timing of `mecat2cns -i 0` at scale: GPU driver vs the unmodified reference binary (all host
cores), on candidates produced by the GPU `mecat2pw -j 0` from synthetic CLR reads (BASELINE configs[3]
shape).  Sorted corrected-FASTA records must be identical.  Writes gpurun_out/fullscale_cns_<reads>.json.
Test/bench tooling (executes oracle/_ref as checker and CPU baseline)."""
import argparse
import hashlib
import json
import os
import subprocess
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def records(path):
    lines = open(path).read().splitlines()
    recs = sorted(zip(lines[0::2], lines[1::2]))
    h = hashlib.sha256()
    for a, b in recs:
        h.update(a.encode()); h.update(b"\n"); h.update(b.encode()); h.update(b"\n")
    return len(recs), sum(len(b) for _, b in recs), h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=20000)
    ap.add_argument("--coverage", type=int, default=15)
    ap.add_argument("--tmp", default="/tmp/mecat_fullscale_cns")
    ap.add_argument("--skip-ref", action="store_true")
    a = ap.parse_args()
    os.makedirs(a.tmp, exist_ok=True)
    fa = os.path.join(a.tmp, "reads.fa")
    genome = a.reads * 15000 // a.coverage
    subprocess.check_call([os.path.join(ROOT, "mecat_b200", "bin", "gen_reads"), fa, str(a.reads), str(genome), "11"])
    can = os.path.join(a.tmp, "cand.can")
    subprocess.check_call("rm -rf %s/w" % a.tmp, shell=True)
    t = time.time()
    subprocess.check_call([os.path.join(ROOT, "mecat_b200", "bin", "mecat2pw"), "-j", "0", "-d", fa, "-o", can, "-w", os.path.join(a.tmp, "w")],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    res = {"reads": a.reads, "genome": genome, "coverage": a.coverage, "cores": os.cpu_count(), "gpu_pw_j0_seconds": time.time() - t,
           "candidates": sum(1 for _ in open(can))}
    gout = os.path.join(a.tmp, "gpu.fa")
    t = time.time()
    with open(os.path.join(a.tmp, "gpu.log"), "w") as lg:
        subprocess.check_call([os.path.join(ROOT, "mecat_b200", "bin", "mecat2cns"), "-i", "0", "-t", "1", can, fa, gout], stdout=lg, stderr=lg,
                              env=dict(os.environ, MECAT_B200_STATS="1"))
    res["gpu_cli_seconds"] = time.time() - t
    res["gpu_log"] = [l for l in open(os.path.join(a.tmp, "gpu.log")).read().splitlines() if "takes" in l or "kernel ms" in l]
    n, bases, sha = records(gout)
    res.update(gpu_records=n, gpu_corrected_bases=bases, gpu_sha256=sha)
    if not a.skip_ref:
        rout = os.path.join(a.tmp, "ref.fa")
        can2 = os.path.join(a.tmp, "cand_ref.can")
        subprocess.check_call(["cp", can, can2])
        t = time.time()
        with open(os.path.join(a.tmp, "ref.log"), "w") as lg:
            subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "mecat2cns"), "-i", "0", "-t", str(os.cpu_count()), can2, fa, rout],
                                  stdout=lg, stderr=lg)
        res["ref_cli_seconds"] = time.time() - t
        res["ref_log"] = [l for l in open(os.path.join(a.tmp, "ref.log")).read().splitlines() if "takes" in l]
        n2, bases2, sha2 = records(rout)
        res.update(ref_records=n2, ref_corrected_bases=bases2, ref_sha256=sha2, identical=(sha == sha2),
                   speedup_cli_wall=res["ref_cli_seconds"] / res["gpu_cli_seconds"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fullscale_cns_%d.json" % a.reads), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

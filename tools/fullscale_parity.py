#!/usr/bin/env python
"""Full-size parity + timing of the two command-line drivers on the same box:

    oracle/_ref/mecat2pw (unmodified reference, -t <all cores>)  vs  mecat_b200/bin/mecat2pw (GPU)

on BASELINE configs[1] (100 000 x 15 kb synthetic CLR reads) or a smaller --reads N.  Sorted outputs
must be byte-identical.  Writes gpurun_out/fullscale_parity.json.  Test/bench tooling (it executes
oracle/_ref as the checker and CPU baseline)."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sorted_sha(path):
    out = path + ".sorted"
    subprocess.check_call("LC_ALL=C sort %s > %s" % (path, out), shell=True)
    h = hashlib.sha256()
    n = 0
    with open(out, "rb") as f:
        for line in f:
            h.update(line)
            n += 1
    return h.hexdigest(), n, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100000)
    ap.add_argument("--job", type=int, default=1)
    ap.add_argument("--tmp", default="/tmp/mecat_fullscale")
    ap.add_argument("--skip-ref", action="store_true")
    a = ap.parse_args()
    os.makedirs(a.tmp, exist_ok=True)
    fa = os.path.join(a.tmp, "reads.fa")
    genome = a.reads * 1000
    subprocess.check_call([os.path.join(ROOT, "mecat_b200", "bin", "gen_reads"), fa, str(a.reads), str(genome), "11"])
    res = {"reads": a.reads, "genome": genome, "seed": 11, "job": a.job, "cores": os.cpu_count()}
    extra = ["-g", "1"] if a.job == 1 else []
    t = time.time()
    gout = os.path.join(a.tmp, "gpu.out")
    subprocess.check_call("rm -rf %s/wg" % a.tmp, shell=True)
    with open(os.path.join(a.tmp, "gpu.log"), "w") as lg:
        subprocess.check_call([os.path.join(ROOT, "mecat_b200", "bin", "mecat2pw"), "-j", str(a.job), "-d", fa, "-o", gout,
                               "-w", os.path.join(a.tmp, "wg"), "-t", "1"] + extra, stdout=lg, stderr=lg)
    res["gpu_cli_seconds"] = time.time() - t
    res["gpu_log"] = [l for l in open(os.path.join(a.tmp, "gpu.log")).read().splitlines() if "takes" in l]
    gsha, gn, gsorted = sorted_sha(gout)
    res["gpu_records"] = gn
    res["gpu_sorted_sha256"] = gsha
    if not a.skip_ref:
        t = time.time()
        rout = os.path.join(a.tmp, "ref.out")
        subprocess.check_call("rm -rf %s/wr" % a.tmp, shell=True)
        with open(os.path.join(a.tmp, "ref.log"), "w") as lg:
            subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "mecat2pw"), "-j", str(a.job), "-d", fa, "-o", rout,
                                   "-w", os.path.join(a.tmp, "wr"), "-t", str(os.cpu_count())] + extra, stdout=lg, stderr=lg)
        res["ref_cli_seconds"] = time.time() - t
        res["ref_log"] = [l for l in open(os.path.join(a.tmp, "ref.log")).read().splitlines() if "takes" in l]
        rsha, rn, rsorted = sorted_sha(rout)
        res["ref_records"] = rn
        res["ref_sorted_sha256"] = rsha
        res["identical"] = (rsha == gsha)
        res["speedup_cli_wall"] = res["ref_cli_seconds"] / res["gpu_cli_seconds"]
        if rsha != gsha:
            subprocess.call("diff %s %s | head -40 > %s" % (rsorted, gsorted, os.path.join(ROOT, "gpurun_out", "fullscale_diff.txt")), shell=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fullscale_parity_%d_j%d.json" % (a.reads, a.job)), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

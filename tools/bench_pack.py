#!/usr/bin/env python
"""2-bit packing of BASELINE configs[1]'s reads (1.6 GB of FASTA): the host packer of the drivers (mecat_b200_split_dataset /
volume_from_fasta, all host threads) next to the device packer (mecat_b200_volume_from_text: H2D of the letters, k_pack_text,
packed bytes back for the volume file).  Writes gpurun_out/bench_pack.json.  Bench tooling."""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mecat_b200  # noqa: E402


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    tmp = "/tmp/bench_pack"
    os.makedirs(tmp, exist_ok=True)
    fa = os.path.join(tmp, "reads_%d.fa" % reads)
    if not os.path.exists(fa):
        subprocess.check_call([os.path.join(ROOT, "mecat_b200", "bin", "gen_reads"), fa, str(reads), str(reads * 1000), "11"])
    res = {"reads": reads, "fasta_bytes": os.path.getsize(fa), "cores": os.cpu_count()}
    t = time.perf_counter()
    hv = mecat_b200.volume_from_fasta(fa)
    res["host_pack_seconds_first"] = time.perf_counter() - t
    t = time.perf_counter()
    hv = mecat_b200.volume_from_fasta(fa)
    res["host_pack_seconds"] = time.perf_counter() - t
    text = open(fa, "rb").read()
    # record boundaries the way the host parser finds them (single-line records of the generator): header line, sequence line
    nl = np.flatnonzero(np.frombuffer(text, dtype=np.uint8) == 10)
    src = nl[0::2] + 1
    lens = nl[1::2] - src
    osz = np.asarray(hv.offset_size).reshape(-1, 2)
    assert len(src) == len(osz) and (lens == osz[:, 1]).all()
    ctx = mecat_b200.Context(0)
    best = None
    for rep in range(3):
        ctx.reset_stats()
        t = time.perf_counter()
        d, pac = ctx.volume_from_text(text, src, osz, hv.num_bases)
        dt = time.perf_counter() - t
        st = ctx.stats()
        ctx.release_volume(d)
        best = dt if best is None else min(best, dt)
        res["device_pack_kernel_ms"] = st["kernel_ms"]["orient"]
    res["device_pack_seconds"] = best
    res["identical_bytes"] = bool((pac == np.frombuffer(bytes(hv.pac[:len(pac)]), dtype=np.uint8)).all())
    res["note"] = ("device_pack_seconds includes the pageable H2D copy of the letters, "
                   "k_pack_text + the two re-layout kernels, and the D2H copy of the packed bytes")
    ctx.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_pack.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

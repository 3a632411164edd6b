#!/bin/bash
# mecat2asmpw at 100 000 reads (400 Mb of corrected-read like letters, 32x): the CUDA path next to the unmodified binary with
# all host threads, records compared line by line.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
timeout 1200 python tools/bench_asm.py --reads 100000 --genome 12500000 --steps 2 > gpurun_out/bench_asm_100k.json 2> gpurun_out/bench_asm_100k.err; tail -c 1800 gpurun_out/bench_asm_100k.json; tail -3 gpurun_out/bench_asm_100k.err

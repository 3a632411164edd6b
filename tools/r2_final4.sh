#!/bin/bash
# Last pass of round 2 at HEAD (compact block records in the mecat2asmpw path): every GPU test, smoke, mecat2asmpw at
# 100 000 reads next to the unmodified binary with the records compared line by line.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/final4_pytest_gpu.log 2>&1; grep -n "passed\|failed" gpurun_out/final4_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python tools/bench_asm.py --reads 100000 --genome 12500000 --steps 2 > gpurun_out/final4_bench_asm_100k.json 2> gpurun_out/final4_bench_asm_100k.err; tail -c 1500 gpurun_out/final4_bench_asm_100k.json; tail -3 gpurun_out/final4_bench_asm_100k.err

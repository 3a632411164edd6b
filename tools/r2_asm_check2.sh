#!/bin/bash
# hardware check of the mecat2asmpw path after the compact block records: GPU tests, ABI bench at 20 000 and 100 000 reads
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_asm_gpu.py -x -q) > gpurun_out/check2_asm_pytest_gpu.log 2>&1; tail -4 gpurun_out/check2_asm_pytest_gpu.log
timeout 600 python tools/bench_asm.py --steps 3 --no-ref > gpurun_out/check2_bench_asm.json 2> gpurun_out/check2_bench_asm.err; tail -c 900 gpurun_out/check2_bench_asm.json
timeout 600 python tools/bench_asm.py --reads 100000 --genome 12500000 --steps 2 --no-ref > gpurun_out/check2_bench_asm_100k.json 2> gpurun_out/check2_bench_asm_100k.err; tail -c 900 gpurun_out/check2_bench_asm_100k.json

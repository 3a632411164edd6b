#!/bin/bash
# quick hardware check of the mecat2asmpw path after a change: its GPU tests and the ABI bench without the reference
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_asm_gpu.py -x -q) > gpurun_out/check_asm_pytest_gpu.log 2>&1; tail -4 gpurun_out/check_asm_pytest_gpu.log
timeout 600 python tools/bench_asm.py --steps 5 --no-ref > gpurun_out/check_bench_asm.json 2> gpurun_out/check_bench_asm.err; tail -c 1000 gpurun_out/check_bench_asm.json

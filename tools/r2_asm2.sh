#!/bin/bash
# Second hardware run of the mecat2asmpw path (warp-per-strand seeding): GPU tests, bench next to the unmodified binary
# (the reference timed once, -O0 build only), launch list.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_asm_gpu.py -x -q -s) > gpurun_out/asm2_pytest_gpu.log 2>&1; tail -5 gpurun_out/asm2_pytest_gpu.log
timeout 900 python tools/bench_asm.py --steps 3 > gpurun_out/bench_asm2.json 2> gpurun_out/bench_asm2.err; tail -c 1600 gpurun_out/bench_asm2.json; tail -3 gpurun_out/bench_asm2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/asm2_launches.csv python tools/bench_asm.py --steps 1 --no-ref > gpurun_out/asm2_ncu.log 2>&1; tail -2 gpurun_out/asm2_ncu.log

#!/bin/bash
# mecat2asmpw path after the warp-per-candidate alignment: GPU tests, bench without the reference, launch list.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_asm_gpu.py -x -q) > gpurun_out/asm3_pytest_gpu.log 2>&1; tail -5 gpurun_out/asm3_pytest_gpu.log
timeout 600 python tools/bench_asm.py --steps 3 --no-ref > gpurun_out/bench_asm3.json 2> gpurun_out/bench_asm3.err; tail -c 900 gpurun_out/bench_asm3.json; tail -3 gpurun_out/bench_asm3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/asm3_launches.csv python tools/bench_asm.py --steps 1 --no-ref > gpurun_out/asm3_ncu.log 2>&1; tail -2 gpurun_out/asm3_ncu.log

#!/bin/bash
# First hardware run of the mecat2asmpw / mecat2trimpw path: its GPU tests, the new -x 1 -i 1 test, the bench next to the
# unmodified binary, and the launch list of the bench.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_asm_gpu.py tests/test_x1_gpu.py::test_nanopore_m4_input_consensus_matches_reference -x -q -s) > gpurun_out/asm_pytest_gpu.log 2>&1; tail -5 gpurun_out/asm_pytest_gpu.log
timeout 900 python tools/bench_asm.py > gpurun_out/bench_asm.json 2> gpurun_out/bench_asm.err; tail -c 1500 gpurun_out/bench_asm.json; tail -3 gpurun_out/bench_asm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/asm_launches.csv python tools/bench_asm.py --steps 1 --no-ref --reads 5000 --genome 600000 > gpurun_out/asm_ncu.log 2>&1; tail -2 gpurun_out/asm_ncu.log

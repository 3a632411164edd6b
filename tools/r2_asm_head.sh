#!/bin/bash
# mecat2asmpw at HEAD (letters upper-cased on the device, 4 GB table budget, several devices in the driver): GPU tests, both
# arms of bench.py --workload asm, the ABI bench, compute-sanitizer memcheck and racecheck of a small run.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_asm_gpu.py -x -q) > gpurun_out/head_asm_pytest_gpu.log 2>&1; tail -4 gpurun_out/head_asm_pytest_gpu.log
timeout 600 python tools/bench_asm.py --steps 3 --no-ref > gpurun_out/head_bench_asm_abi.json 2> gpurun_out/head_bench_asm_abi.err; tail -c 900 gpurun_out/head_bench_asm_abi.json
timeout 600 python bench.py --workload asm --steps 3 --warmup 1 > gpurun_out/head_bench_asm_ours.json 2> gpurun_out/head_bench_asm_ours.err; tail -c 700 gpurun_out/head_bench_asm_ours.json; tail -4 gpurun_out/head_bench_asm_ours.err
timeout 600 python bench.py --workload asm --impl reference --steps 1 --warmup 0 > gpurun_out/head_bench_asm_reference.json 2> gpurun_out/head_bench_asm_reference.err; tail -c 500 gpurun_out/head_bench_asm_reference.json
(timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/asm_sanitize.py 300; echo rc=$?) > gpurun_out/head_asm_memcheck.log 2>&1; tail -6 gpurun_out/head_asm_memcheck.log
(timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/asm_sanitize.py 60; echo rc=$?) > gpurun_out/head_asm_racecheck.log 2>&1; tail -6 gpurun_out/head_asm_racecheck.log

#!/bin/bash
# Final pass of round 2 at HEAD after the mecat2asmpw row: every GPU test, smoke, the headline bench, the asm workload.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/final3_pytest_gpu.log 2>&1; tail -4 gpurun_out/final3_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/final3_bench_n1.json 2> gpurun_out/final3_bench_n1.err; tail -c 600 gpurun_out/final3_bench_n1.json
python bench.py --workload asm --steps 3 --warmup 1 > gpurun_out/final3_asm_ours.json 2> gpurun_out/final3_asm_ours.err; tail -c 500 gpurun_out/final3_asm_ours.json; tail -4 gpurun_out/final3_asm_ours.err

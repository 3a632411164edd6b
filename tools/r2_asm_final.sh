#!/bin/bash
# Evidence of the mecat2asmpw / mecat2trimpw path at HEAD: GPU tests, bench next to the unmodified binary (as built and
# with -O2), launch list, one full ncu capture of the seeding and of the alignment kernel (5 000-read input).
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_asm_gpu.py -x -q) > gpurun_out/asm_final_pytest_gpu.log 2>&1; tail -5 gpurun_out/asm_final_pytest_gpu.log
timeout 900 python tools/bench_asm.py --steps 3 > gpurun_out/bench_asm_final.json 2> gpurun_out/bench_asm_final.err; tail -c 1800 gpurun_out/bench_asm_final.json; tail -3 gpurun_out/bench_asm_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/asm_final_launches.csv python tools/bench_asm.py --steps 1 --no-ref > gpurun_out/asm_final_ncu.log 2>&1; tail -2 gpurun_out/asm_final_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_asm_seed_warp|k_asm_slots' -c 2 -f -o gpurun_out/asm_final_full python tools/bench_asm.py --steps 0 --no-ref --reads 5000 --genome 600000 > gpurun_out/asm_final_ncu_full.log 2>&1; tail -2 gpurun_out/asm_final_ncu_full.log
ls -la gpurun_out/asm_final_full.ncu-rep

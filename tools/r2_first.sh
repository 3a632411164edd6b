#!/bin/bash
# Round-2 first hardware pass: the GPU tests that never ran on hardware first, then the rest, the headline bench at HEAD,
# fresh ncu captures of k_seed / k_extend, and the first mecat2ref measurement at BASELINE configs[2] size.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv; nproc; free -g | head -2; df -h /tmp | tail -1
(time timeout 1500 python -m pytest tests/test_ref_gpu.py tests/test_tiles_cli_gpu.py -m gpu -q) > gpurun_out/r2_pytest_ref.log 2>&1; tail -25 gpurun_out/r2_pytest_ref.log
(time timeout 1500 python -m pytest tests/test_gpu.py -m gpu -q) > gpurun_out/r2_pytest_pw.log 2>&1; tail -8 gpurun_out/r2_pytest_pw.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_head.json 2> gpurun_out/r2_bench_n1_head.err; tail -c 600 gpurun_out/r2_bench_n1_head.json
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_head.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_seed|k_extend$' -c 6 -f -o gpurun_out/r2_seed_extend_head python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2_ncu_seed_extend.log 2>&1; tail -2 gpurun_out/r2_ncu_seed_extend.log
timeout 1500 python tools/bench_ref.py --reads 100000 --sample 8000 > gpurun_out/r2_bench_ref.log 2>&1; tail -40 gpurun_out/r2_bench_ref.log
ls -la gpurun_out | tail -12

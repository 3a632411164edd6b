#!/bin/bash
# One GPU-box pass that refreshes the evidence under gpurun_out/ (copied into profiles/ afterwards):
# GPU tests, headline bench (both arms), ncu launch list + full capture of the dominant kernel of the bench,
# ncu launch list + full capture of the mecat2cns kernels, command-line timings at 100 000 reads, the mecat2ref pass
# (tools/profile_ref.sh).
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
(time python -m pytest tests -m gpu -x -q) > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 400 gpurun_out/final_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -c 300 gpurun_out/final_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_extend$' -c 8 -f -o gpurun_out/final_extend_full python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/final_ncu_extend.log 2>&1; tail -2 gpurun_out/final_ncu_extend.log
bash tools/profile_cns.sh 4000
python tools/fullscale_cns.py --reads 100000 --skip-ref > gpurun_out/final_cns100k.log 2>&1; grep -E "kernel ms|takes|seconds|sha|records" gpurun_out/final_cns100k.log
python tools/fullscale_parity.py --skip-ref > gpurun_out/final_pw100k.log 2>&1; grep -E "takes|seconds|sha|records" gpurun_out/final_pw100k.log
bash tools/profile_ref.sh 20000 2000
ls -la gpurun_out | tail -20

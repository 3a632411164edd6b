#!/bin/bash
# Refresh of the evidence that changed after tools/final_evidence.sh ran (k_xdrop diet, mecat2cns stack fix, -i 1 test):
# all GPU tests, the -x 1 bench (ours) with a fresh full capture of k_xdrop, the cns / ref workloads, smoke.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --tech 1 --steps 3 --warmup 1 --no-cpu > gpurun_out/final_x1_ours.json 2> gpurun_out/final_x1_ours.err; tail -c 300 gpurun_out/final_x1_ours.json
ncu --set full --clock-control none --import-source on -k regex:'k_xdrop' -c 1 -f -o gpurun_out/final_xdrop_full python bench.py --tech 1 --steps 1 --warmup 0 --no-cpu > gpurun_out/final_ncu_xdrop.log 2>&1; tail -2 gpurun_out/final_ncu_xdrop.log
for w in ref cns; do
  python bench.py --workload $w --steps 5 --warmup 1 > gpurun_out/final_${w}_ours.json 2> gpurun_out/final_${w}_ours.err; tail -c 300 gpurun_out/final_${w}_ours.json
done
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/final_bench_n1_nocpu.json 2> /dev/null; tail -c 200 gpurun_out/final_bench_n1_nocpu.json

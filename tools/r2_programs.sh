#!/bin/bash
# BASELINE configs[2] and configs[3] through bench.py (command-line drivers), both arms.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
for w in ref cns; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 1 > gpurun_out/r2p_${w}_ours.json 2> gpurun_out/r2p_${w}_ours.err; tail -c 700 gpurun_out/r2p_${w}_ours.json; tail -2 gpurun_out/r2p_${w}_ours.err
  timeout 1200 python bench.py --workload $w --impl reference --steps 1 --warmup 0 > gpurun_out/r2p_${w}_reference.json 2> gpurun_out/r2p_${w}_reference.err; tail -c 900 gpurun_out/r2p_${w}_reference.json; tail -2 gpurun_out/r2p_${w}_reference.err
done

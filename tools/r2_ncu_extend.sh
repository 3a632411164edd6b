#!/bin/bash
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend' -c 2 -f -o gpurun_out/r2_extend_head2 python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2_ncu_extend2.log 2>&1; tail -2 gpurun_out/r2_ncu_extend2.log

#!/bin/bash
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
for cfg in ${CFGS:-"pairs2 8" "pairs3 8" "pairs3 1"}; do
  set -- $cfg
  MECAT_B200_EXTEND=$1 MECAT_B200_PIPE=$2 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/r2x_bench_$1_$2.json 2> gpurun_out/r2x_bench_$1_$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2x_bench_$1_$2.json"))
print("$1 pipe $2", d["ms_per_step"], d["pairs_per_step"], d["kernel_ms_per_step"]["extend"], d["deterministic"], d["wall_ms_per_step"])
PY
done
if [ -n "$NCU" ]; then
MECAT_B200_EXTEND=$NCU MECAT_B200_PIPE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend_pairs' -c 1 -f -o gpurun_out/r2x_$NCU python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2x_ncu_$NCU.log 2>&1; tail -2 gpurun_out/r2x_ncu_$NCU.log
fi

#!/bin/bash
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu.py -m gpu -q -x) > gpurun_out/r2s_pytest_pw.log 2>&1; tail -8 gpurun_out/r2s_pytest_pw.log
for ctas in 2 1; do
  MECAT_B200_SEED_CTAS=$ctas timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r2s_bench_ctas$ctas.json 2> gpurun_out/r2s_bench_ctas$ctas.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2s_bench_ctas$ctas.json"))
print("ctas $ctas", d["ms_per_step"], d["pairs_per_step"], d["kernel_ms_per_step"]["seed"], d["deterministic"], d["kernel_ms_per_step"])
PY
done

#!/bin/bash
# A/B of the extension kernels: parity tests, then the headline bench per kernel form.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -x) > gpurun_out/r2x_pytest_pw.log 2>&1; tail -8 gpurun_out/r2x_pytest_pw.log
for mode in ${MODES:-pairs2 pairs3 pairs1 warp}; do
  MECAT_B200_EXTEND=$mode timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/r2x_bench_$mode.json 2> gpurun_out/r2x_bench_$mode.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2x_bench_$mode.json"))
print("$mode", d["ms_per_step"], d["pairs_per_step"], d["kernel_ms_per_step"]["extend"], d["deterministic"], d.get("roofline",{}).get("issue"))
PY
done

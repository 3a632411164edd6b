#!/bin/bash
# quick pass: extension / alignment parity tests + headline bench (no CPU leg)
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -x -k "${TESTS:-extend or tiles or cfg0 or small}") > gpurun_out/r2q_pytest.log 2>&1; tail -5 gpurun_out/r2q_pytest.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2q_bench.json"))
print("bench", d["ms_per_step"], d["pairs_per_step"], d["deterministic"], {k:v for k,v in d["kernel_ms_per_step"].items() if v}, d["roofline"].get("issue",{}).get("cells_per_s"))
PY

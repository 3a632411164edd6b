#!/bin/bash
# Single-GPU side of the configs[4] scaling figures: GPU tests touched since the last pass, ring-mode bench at N = 1 (all 36
# tiles on one device), the command-line driver on the same 1 M reads with one device.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_ref_gpu.py::test_degenerate_inputs tests/test_gpu.py::test_repeat_heavy_long_reads tests/test_gpu.py::test_extend_matches_oracle -m gpu -q) > gpurun_out/r2_pytest_touch.log 2>&1; tail -6 gpurun_out/r2_pytest_touch.log
timeout 900 python bench.py --gpus 1 --mode ring --volumes 8 --steps 1 --warmup 1 > gpurun_out/r2m_ring_n1.json 2> gpurun_out/r2m_ring_n1.err; tail -c 900 gpurun_out/r2m_ring_n1.json; tail -3 gpurun_out/r2m_ring_n1.err
bash tools/r2_cfg4.sh 1

#!/bin/bash
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --mode ring --volumes 8 --steps 2 --warmup 1 > gpurun_out/r2m_ring_n$N.json 2> gpurun_out/r2m_ring_n$N.err; tail -c 1200 gpurun_out/r2m_ring_n$N.json; tail -3 gpurun_out/r2m_ring_n$N.err
MECAT_B200_SPLIT_TIMING=1 bash tools/r2_cfg4.sh 8
grep -E "\[split\]|takes" /tmp/cfg4/cli_n8.err | head -5

#!/bin/bash
# Multi-GPU pass: N = $1 GPUs.  strong mode (configs[1] tile shared) and ring mode (configs[4] job) through bench.py.
set -x
N=${1:-2}
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-3} --warmup 2 > gpurun_out/r2m_strong_n$N.json 2> gpurun_out/r2m_strong_n$N.err; tail -c 1500 gpurun_out/r2m_strong_n$N.json; tail -5 gpurun_out/r2m_strong_n$N.err
timeout 900 $TR bench.py --gpus $N --mode ring ${RINGARGS:---volumes 8} --steps ${RSTEPS:-1} --warmup 1 > gpurun_out/r2m_ring_n$N.json 2> gpurun_out/r2m_ring_n$N.err; tail -c 1500 gpurun_out/r2m_ring_n$N.json; tail -5 gpurun_out/r2m_ring_n$N.err

#!/bin/bash
# First hardware pass for the mecat2ref row (run on the GPU box through gpurun; writes under gpurun_out/):
#   ref_pytest_gpu.log     tests/test_ref_gpu.py (parity against the unmodified binary's golden files)
#   bench_ref_<reads>.json parity on a sample + timing against the unmodified binary (tools/bench_ref.py)
#   ref_launches.csv       per-launch durations of one mecat2ref run
#   ref_full.ncu-rep       --set full capture of k_ref (count, seed, rescue) and k_align of the same command
# usage: bash tools/profile_ref.sh [reads=20000] [sample=2000]
set -x
READS=${1:-20000}
SAMPLE=${2:-2000}
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_ref_gpu.py -m gpu -q) > gpurun_out/ref_pytest_gpu.log 2>&1; tail -5 gpurun_out/ref_pytest_gpu.log
timeout 1200 python tools/bench_ref.py --reads $READS --sample $SAMPLE > gpurun_out/bench_ref.log 2>&1; tail -30 gpurun_out/bench_ref.log
cp gpurun_out/bench_ref_$READS.json gpurun_out/bench_ref_${READS}_strings_kernel.json
timeout 600 python tools/bench_ref.py --reads $READS --sample $SAMPLE --forward --skip-ref > gpurun_out/bench_ref_forward.log 2>&1; tail -12 gpurun_out/bench_ref_forward.log
cp gpurun_out/bench_ref_$READS.json gpurun_out/bench_ref_${READS}_forward.json
cd /tmp/mecat_bench_ref
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $ROOT/gpurun_out/ref_launches.csv \
    $ROOT/mecat_b200/bin/mecat2ref -d reads.fa -r genome.fa -o ncu1.m4 -w wn1 -m 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ref|k_align' -c 12 -f -o $ROOT/gpurun_out/ref_full \
    $ROOT/mecat_b200/bin/mecat2ref -d reads.fa -r genome.fa -o ncu2.m4 -w wn2 -m 1 > $ROOT/gpurun_out/ncu_ref_full.log 2>&1
tail -3 $ROOT/gpurun_out/ncu_ref_full.log

#!/bin/bash
# ncu --set full of the mecat2ref seeding kernel in both shapes (warp per strand = default, thread per strand) on BASELINE configs[2]
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT; mkdir -p gpurun_out
T=/tmp/ncu_refseed; mkdir -p $T
B=$ROOT/mecat_b200/bin
$B/gen_reads $T/reads.fa 100000 100000000 11 15000 1500 0.15 $T/genome.fa
cd $T
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ref_seed_warp -c 1 -f -o $ROOT/gpurun_out/r2_refseed_warp \
    $B/mecat2ref -d reads.fa -r genome.fa -o o1.m4 -w w1 -m 1 > $ROOT/gpurun_out/r2_ncu_refseed_warp.log 2>&1
tail -n 2 $ROOT/gpurun_out/r2_ncu_refseed_warp.log

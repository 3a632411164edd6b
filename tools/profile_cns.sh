#!/bin/bash
# ncu evidence for the mecat2cns kernels (run on the GPU box through gpurun; writes under gpurun_out/):
#   cns_launches.csv       per-launch durations of one mecat2cns run (4 000 reads)
#   cns_full.ncu-rep       --set full capture of the consensus kernels and k_align of the same command
# usage: bash tools/profile_cns.sh [reads]
set -x
READS=${1:-4000}
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
python $ROOT/tools/fullscale_cns.py --reads $READS --skip-ref > /dev/null 2>&1
cd /tmp/mecat_fullscale_cns
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $ROOT/gpurun_out/cns_launches.csv \
    $ROOT/mecat_b200/bin/mecat2cns -i 0 -t 1 cand.can reads.fa ncu1.fa > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_cns$|k_cns_warp|k_cns_poa|k_align' -f -o $ROOT/gpurun_out/cns_full \
    $ROOT/mecat_b200/bin/mecat2cns -i 0 -t 1 cand.can reads.fa ncu2.fa > $ROOT/gpurun_out/ncu_cns_full.log 2>&1
tail -3 $ROOT/gpurun_out/ncu_cns_full.log

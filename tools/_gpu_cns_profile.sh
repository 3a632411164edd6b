set -x
python -m pytest tests -m gpu -x -q -k cns 2>&1 | tail -3
python tools/fullscale_cns.py --reads 20000 --skip-ref > gpurun_out/cns20k_c.log 2>&1; grep -E "kernel ms|takes|seconds|sha" gpurun_out/cns20k_c.log
python tools/fullscale_cns.py --reads 4000 --skip-ref > /dev/null 2>&1
cd /tmp/mecat_fullscale_cns
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/cns_launches.csv $GRAFT_REPO_ROOT/mecat_b200/bin/mecat2cns -i 0 -t 1 cand.can reads.fa ncu1.fa > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_cns$|k_cns_warp' -o $GRAFT_REPO_ROOT/gpurun_out/cns_full $GRAFT_REPO_ROOT/mecat_b200/bin/mecat2cns -i 0 -t 1 cand.can reads.fa ncu2.fa > $GRAFT_REPO_ROOT/gpurun_out/ncu_full.log 2>&1
tail -3 $GRAFT_REPO_ROOT/gpurun_out/ncu_full.log; ls -la $GRAFT_REPO_ROOT/gpurun_out/

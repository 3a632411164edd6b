#!/bin/bash
# BASELINE configs[4] through the command-line driver: 1 M x 15 kb reads (genome 1 Gb, seed 13), the splitter's own
# volumes (7 full + a sliver), MECAT_GPUS=$1 devices.  Writes gpurun_out/r2_cfg4_cli_n$1.json.
set -x
N=${1:-1}
READS=${2:-1000000}
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out /tmp/cfg4
FA=/tmp/cfg4/reads_$READS.fa
if [ ! -s $FA ]; then ( time mecat_b200/bin/gen_reads $FA $READS $((READS * 1000)) 13 ) 2>&1 | tail -3; fi
ls -la $FA
rm -rf /tmp/cfg4/w$N
T0=$(date +%s.%N)
MECAT_GPUS=$N mecat_b200/bin/mecat2pw -j 1 -d $FA -o /tmp/cfg4/out_n$N.m4 -w /tmp/cfg4/w$N -t 16 > /tmp/cfg4/cli_n$N.out 2> /tmp/cfg4/cli_n$N.err
RC=$?
T1=$(date +%s.%N)
grep -E "split_raw_dataset\] takes|merge_results\] takes" /tmp/cfg4/cli_n$N.err
python - <<PY
import json, re, subprocess, sys
sys.path.insert(0, "tools")
from lines_digest import digest
err = open("/tmp/cfg4/cli_n$N.err").read()
tiles = [float(x) for x in re.findall(r"\[process volume \d+\] takes ([0-9.]+) secs", err)]
idx = [float(x) for x in re.findall(r"\[create_ref_index\] takes ([0-9.]+) secs", err)]
split = [float(x) for x in re.findall(r"\[split_raw_dataset\] takes ([0-9.]+) secs", err)]
merge = [float(x) for x in re.findall(r"\[merge_results\] takes ([0-9.]+) secs", err)]
wall = $T1 - $T0
d = digest("/tmp/cfg4/out_n$N.m4") if $RC == 0 else {}
vols = len(open("/tmp/cfg4/w$N/fileindex.txt").read().split())
res = {"reads": $READS, "gpus": $N, "rc": $RC, "volumes": vols, "tiles": len(tiles), "cli_wall_seconds": wall,
       "split_seconds": split, "merge_seconds": merge, "sum_tile_seconds": sum(tiles), "index_builds": len(idx), "sum_index_seconds": sum(idx),
       "tile_phase_seconds": wall - sum(split) - sum(merge), "records": d.get("lines"), "digest": d,
       "pairs_per_second_cli": (d.get("lines") or 0) / wall, "pairs_per_second_tile_phase": (d.get("lines") or 0) / max(1e-9, wall - sum(split) - sum(merge))}
json.dump(res, open("gpurun_out/r2_cfg4_cli_n$N.json", "w"), indent=1)
print(json.dumps(res))
PY
cp /tmp/cfg4/cli_n$N.err gpurun_out/r2_cfg4_cli_n$N.err; tail -3 /tmp/cfg4/cli_n$N.err

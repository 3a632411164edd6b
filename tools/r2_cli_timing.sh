#!/bin/bash
# wall clock of the mecat2pw command line on BASELINE configs[1] under the result-path variants (device text, host text,
# host record assembly); reads are generated once
set -x
T=${TMPDIR:-/tmp}/cli_timing; mkdir -p $T gpurun_out
B=mecat_b200/bin
[ -f $T/reads.fa ] || $B/gen_reads $T/reads.fa 100000 100000000 11
for rep in 1 2; do
for v in "device" "TEXT=host" "M4=host TEXT=host"; do
  rm -rf $T/w
  s=$(date +%s.%N)
  if [ "$v" = "device" ]; then $B/mecat2pw -j 1 -d $T/reads.fa -o $T/out.m4 -w $T/w -t 16 2> $T/log.txt
  elif [ "$v" = "TEXT=host" ]; then MECAT_B200_TEXT=host $B/mecat2pw -j 1 -d $T/reads.fa -o $T/out.m4 -w $T/w -t 16 2> $T/log.txt
  else MECAT_B200_M4=host MECAT_B200_TEXT=host $B/mecat2pw -j 1 -d $T/reads.fa -o $T/out.m4 -w $T/w -t 16 2> $T/log.txt; fi
  e=$(date +%s.%N)
  echo "variant [$v] rep $rep: $(python -c "print(round($e - $s, 2))") s; $(grep takes $T/log.txt | tr '\n' ' ')" | tee -a gpurun_out/r2_cli_timing.txt
done
done

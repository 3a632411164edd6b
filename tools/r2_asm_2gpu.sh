#!/bin/bash
# mecat2asmpw on two devices (MECAT_GPUS=2: an index replica per device, the reads of the query file split between them):
# the command-line GPU test (it runs the two-device case when two devices are present) and the bench block file with one
# and with two devices, results compared.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
nvidia-smi -L
(time timeout 600 python -m pytest tests/test_asm_gpu.py -x -q -k command_line) > gpurun_out/asm_2gpu_pytest.log 2>&1; tail -4 gpurun_out/asm_2gpu_pytest.log
W=/tmp/asm_blocks; rm -rf $W; mkdir -p $W
mecat_b200/bin/gen_reads $W/000001.fasta 40000 5000000 11 4000 800 0.015 - 0 > /dev/null
echo "-allreads -allbases -b 1 -e 40000" > $W/ovlprep
{
for g in 1 2 1 2 1 2; do
  rm -f $W/*.r
  t0=$(date +%s%N)
  MECAT_GPUS=$g mecat_b200/bin/mecat2asmpw -P$W -T2 -S1 -E1
  echo "MECAT_GPUS=$g wall $(( ($(date +%s%N) - t0) / 1000000 )) ms"
  echo "MECAT_GPUS=$g sorted md5 $(cat $W/*.r | sort | md5sum | cut -c1-32) lines $(cat $W/*.r | wc -l)"
done
} > gpurun_out/asm_2gpu_cli.log 2>&1
cat gpurun_out/asm_2gpu_cli.log

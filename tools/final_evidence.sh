#!/bin/bash
# One GPU-box pass that refreshes the round's evidence under gpurun_out/ (copied into profiles/ afterwards):
# all GPU tests, headline bench (both arms, driver-style flags), ncu launch list with instruction counts, one
# `--set full` capture of k_seed + k_extend of the bench command, the two other programs (configs[2], [3]) through bench.py,
# command-line runs at 100 000 reads with full-size parity against the committed reference digests.
set -x
ROOT=${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}
cd $ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv; nproc
(time python -m pytest tests -m gpu -x -q) > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 300 gpurun_out/final_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -c 300 gpurun_out/final_bench_reference.json
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
# (gpurun brings back at most 64 MiB: one launch per kernel, ~8 MB each with sources)
ncu --set full --clock-control none --import-source on -k regex:'k_seed' -c 1 -f -o gpurun_out/final_seed_full python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/final_ncu_seed.log 2>&1; tail -2 gpurun_out/final_ncu_seed.log
ncu --set full --clock-control none --import-source on -k regex:'k_extend' -c 1 -f -o gpurun_out/final_extend_full python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/final_ncu_extend.log 2>&1; tail -2 gpurun_out/final_ncu_extend.log
# the nanopore parameter set (-x 1) through the same C-ABI calls, both arms, and one full capture of its extension kernel
python bench.py --tech 1 --steps 3 --warmup 1 --no-cpu > gpurun_out/final_x1_ours.json 2> gpurun_out/final_x1_ours.err; tail -c 300 gpurun_out/final_x1_ours.json
python bench.py --tech 1 --impl reference --steps 1 --warmup 0 > gpurun_out/final_x1_reference.json 2> gpurun_out/final_x1_reference.err; tail -c 300 gpurun_out/final_x1_reference.json
ncu --set full --clock-control none --import-source on -k regex:'k_xdrop' -c 1 -f -o gpurun_out/final_xdrop_full python bench.py --tech 1 --steps 1 --warmup 0 --no-cpu > gpurun_out/final_ncu_xdrop.log 2>&1; tail -2 gpurun_out/final_ncu_xdrop.log
for w in ref cns; do
  python bench.py --workload $w --steps 3 --warmup 1 > gpurun_out/final_${w}_ours.json 2> gpurun_out/final_${w}_ours.err; tail -c 300 gpurun_out/final_${w}_ours.json
  python bench.py --workload $w --impl reference --steps 1 --warmup 0 > gpurun_out/final_${w}_reference.json 2> gpurun_out/final_${w}_reference.err; tail -c 300 gpurun_out/final_${w}_reference.json
done
python tools/fullscale_parity.py --skip-ref > gpurun_out/final_pw100k.log 2>&1; grep -E "takes|seconds|sha|records" gpurun_out/final_pw100k.log
python tools/fullscale_cns.py --reads 100000 --skip-ref > gpurun_out/final_cns100k.log 2>&1; grep -E "kernel ms|takes|seconds|sha|records" gpurun_out/final_cns100k.log
# memory and race checks of the kernels on the small fixtures (slow under the tool: bounded)
(time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu.py -m gpu -q -x -k "small_m4 or small_can or cns_small or extend_matches or raw_candidates or index_matches") > gpurun_out/final_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/final_sanitizer_memcheck.log
(time timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu.py -m gpu -q -x -k "small_m4 or extend_matches") > gpurun_out/final_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/final_sanitizer_racecheck.log
(time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_x1_gpu.py tests/test_records_gpu.py -m gpu -q -x -k "string_free or small_nanopore or tile_text or extreme or packing") > gpurun_out/final_sanitizer_memcheck_r2rows.log 2>&1; echo "memcheck (x1 / records) rc=$?"; tail -4 gpurun_out/final_sanitizer_memcheck_r2rows.log
# the seeding kernel of mecat2ref (warp per strand) on configs[2]
bash tools/r2_ncu_refseed.sh > gpurun_out/final_ncu_refseed.log 2>&1; tail -3 gpurun_out/final_ncu_refseed.log
rm -f gpurun_out/r2_refseed_thread.ncu-rep
rm -f gpurun_out/*.tmp
ls -la gpurun_out | tail -25

#!/bin/bash
# One GPU visit: parity suite, contract bench, launch list and full ncu captures of the hot kernels.
# usage: bash tools/gpu_round.sh <tag> [ncu-kernel-regex]
# NOTE: gpurun returns at most 64 MiB of gpurun_out/: keep the full capture to ~24 launches (a 40-launch capture with
# sources was 70+ MiB and the whole visit's output was dropped in r1q).
TAG=${1:-rX}
KRE=${2:-'clahe|resize_|reduce_kernel|gem_pool|whiten_tc|score_filter|topk_finalize'}
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest_gpu exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/quick_bench.py topk > gpurun_out/quick_bench_topk_$TAG.log 2>&1; echo "quick_topk exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/prof_target.py > gpurun_out/prof_plain.log 2>&1
rc=$?; echo "prof_plain exit $rc" >> gpurun_out/summary.txt
if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/prof_target.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu_launches exit $?" >> gpurun_out/summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -c 24 -o gpurun_out/prof_$TAG -f python tools/prof_target.py > gpurun_out/ncu_full.log 2>&1; echo "ncu_full exit $?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/smoke_$TAG.log
tail -n 25 gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/bench_$TAG.json
tail -n 5 gpurun_out/bench_$TAG.err gpurun_out/quick_bench_topk_$TAG.log gpurun_out/prof_plain.log

#!/bin/bash
# round 2, visit q: config 1 at its stated size, ncu of K4's rank_counts, smoke + whole GPU suite + contract bench
TAG=r2q
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary_$TAG.txt
timeout 1800 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest_gpu exit $?" >> gpurun_out/summary_$TAG.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> gpurun_out/summary_$TAG.txt
timeout 900 python tools/config1_full.py > gpurun_out/config1_full_$TAG.json 2> gpurun_out/config1_full_$TAG.err; echo "config1 exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python tools/prof_target.py map > gpurun_out/prof_plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rank_counts_kernel' -s 1 -c 1 -o gpurun_out/prof_k4_$TAG python tools/prof_target.py map > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary_$TAG.txt
cat gpurun_out/summary_$TAG.txt; tail -n 2 gpurun_out/smoke_$TAG.log; grep -E "passed|failed|FAILED|skipped" gpurun_out/pytest_gpu_$TAG.log | tail -8 | cut -c1-220; tail -n 3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json | cut -c1-300; cat gpurun_out/config1_full_$TAG.json; tail -5 gpurun_out/config1_full_$TAG.err

#!/bin/bash
# round 2, visit an: smoke + the WHOLE GPU suite + contract bench (own arm and reference arm) on the current binary, launch list
TAG=r2an
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary_$TAG.txt
timeout 1800 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest_gpu exit $?" >> gpurun_out/summary_$TAG.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> gpurun_out/summary_$TAG.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "bench ref exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python tools/prof_target.py > gpurun_out/prof_plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/prof_target.py > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?" >> gpurun_out/summary_$TAG.txt
cat gpurun_out/summary_$TAG.txt; tail -n 2 gpurun_out/smoke_$TAG.log; grep -E "passed|failed|FAILED|skipped" gpurun_out/pytest_gpu_$TAG.log | tail -8 | cut -c1-220; tail -n 3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json | cut -c1-300; cut -c1-500 gpurun_out/bench_${TAG}_ref.json

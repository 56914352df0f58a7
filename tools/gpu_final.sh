#!/bin/bash
# final confirmation of the round's binary without profilers: smoke, whole GPU suite, contract bench
TAG=${1:-rX}
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest_gpu exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 2 gpurun_out/smoke_$TAG.log; tail -n 4 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/bench_$TAG.json; tail -n 5 gpurun_out/bench_$TAG.err

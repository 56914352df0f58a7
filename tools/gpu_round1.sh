#!/bin/bash
# first GPU visit: parity tests file by file (each under its own timeout), then quick timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for t in test_gpu_clahe test_gpu_gem test_gpu_map; do
  timeout 300 python -m pytest tests/$t.py -q -m gpu -x --timeout 120 > gpurun_out/$t.log 2>&1; echo "$t exit $?" >> gpurun_out/summary.txt
done
timeout 300 python -m pytest tests/test_gpu_topk.py -q -m gpu -k "exact_path or merge or golden" --timeout 120 > gpurun_out/test_gpu_topk_exact.log 2>&1; echo "topk_exact exit $?" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_topk.py -q -m gpu -k "tcgen05" --timeout 100 > gpurun_out/test_gpu_topk_tc.log 2>&1; echo "topk_tc exit $?" >> gpurun_out/summary.txt
timeout 200 python tools/quick_bench.py clahe gem > gpurun_out/quick_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
if grep -q "topk_tc exit 0" gpurun_out/summary.txt; then
  timeout 300 python tools/quick_bench.py topk > gpurun_out/quick_bench_topk.log 2>&1; echo "bench_topk exit $?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
tail -n 30 gpurun_out/*.log

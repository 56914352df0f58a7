#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/k3_ab.log
for c in 1 2 4; do
  GDT_DEBUG_K3_CLUSTER=$c timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_map.py -q -m gpu --timeout 200 -x > gpurun_out/pytest_k3_c$c.log 2>&1; echo "pytest cluster=$c exit $?" >> gpurun_out/summary.txt
  echo "== cluster $c" >> gpurun_out/k3_ab.log
  GDT_DEBUG_K3_CLUSTER=$c timeout 300 python tools/quick_bench.py topk >> gpurun_out/k3_ab.log 2>&1
done
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_k3_c*.log; cat gpurun_out/k3_ab.log

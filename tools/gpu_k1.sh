#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/k1_ab.log
for t in 0 1 2; do
  echo "== splmode $t" >> gpurun_out/k1_ab.log
  GDT_DEBUG_K1_SPL=$t timeout 200 python tools/k1_ab.py >> gpurun_out/k1_ab.log 2>&1
  GDT_DEBUG_K1_SPL=$t timeout 600 python -m pytest tests/test_gpu_clahe.py -q -m gpu --timeout 300 > gpurun_out/pytest_k1_$t.log 2>&1; echo "pytest spl=$t exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; cat gpurun_out/k1_ab.log

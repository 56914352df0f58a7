#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/k1_ab.log
for t in 0 2 8 18 24 16; do
  echo "== texmode $t" >> gpurun_out/k1_ab.log
  GDT_DEBUG_K1_TEX=$t timeout 200 python tools/k1_ab.py >> gpurun_out/k1_ab.log 2>&1
done
GDT_DEBUG_K1_TEX=18 timeout 600 python -m pytest tests/test_gpu_clahe.py -q -m gpu --timeout 300 > gpurun_out/pytest_k1.log 2>&1; echo "pytest tex=18 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_k1.log; cat gpurun_out/k1_ab.log

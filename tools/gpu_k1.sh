#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/k1_ab.log
timeout 600 python -m pytest tests/test_gpu_clahe.py -q -m gpu --timeout 300 > gpurun_out/pytest_k1.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
for o in 4 6 8; do GDT_DEBUG_K1_OCC=$o timeout 200 python tools/k1_ab.py >> gpurun_out/k1_ab.log 2>&1; done
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_k1.log; cat gpurun_out/k1_ab.log

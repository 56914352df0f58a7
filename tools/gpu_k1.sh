#!/bin/bash
# K1 / K5 visit: bit-exactness suites (pipeline suite twice: race detector) 
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/k1_ab.log
timeout 600 python -m pytest tests/test_gpu_resize.py tests/test_gpu_pipeline.py -q -m gpu --timeout 300 > gpurun_out/pytest_k1.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 300 > gpurun_out/pytest_k1b.log 2>&1; echo "pytest2 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -30 gpurun_out/pytest_k1.log; tail -5 gpurun_out/pytest_k1b.log

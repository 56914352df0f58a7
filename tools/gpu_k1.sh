#!/bin/bash
# K1 / K5 visit: bit-exactness suites + K1 timing (default config, rows sweep, odd sizes)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/k1_ab.log
timeout 600 python -m pytest tests/test_gpu_clahe.py tests/test_gpu_hub.py tests/test_gpu_pipeline.py -q -m gpu --timeout 300 > gpurun_out/pytest_k1.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/k1_ab.py >> gpurun_out/k1_ab.log 2>&1
cat gpurun_out/summary.txt; grep -v "Warning\|fork\|^$\|outs =" gpurun_out/pytest_k1.log | tail -8 | cut -c1-250; grep -v "^rows_per_cta=[1-9]" gpurun_out/k1_ab.log

#!/bin/bash
# pipeline suite (device-side geometry is the default) + resize suite
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_resize.py tests/test_gpu_hub.py -q -m gpu --timeout 300 > gpurun_out/pytest_k1.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -v "Warning\|fork\|^$" gpurun_out/pytest_k1.log | tail -12 | cut -c1-300

#!/bin/bash
# race hunt: the pipeline suite four times
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for i in 1 2 3 4; do
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 300 > gpurun_out/pytest_k1_$i.log 2>&1; echo "pytest $i exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; grep -h "^E  \|passed\|failed" gpurun_out/pytest_k1_*.log | cut -c1-300

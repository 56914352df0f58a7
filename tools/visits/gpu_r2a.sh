#!/bin/bash
# round 2, visit a: new wrapper / extraction tests + K1 chunk A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hub.py tests/test_gpu_pipeline.py tests/test_gpu_clahe.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_r2a.log 2>&1; echo "pytest exit $?" > gpurun_out/summary_r2a.txt
timeout 300 python tools/k1_chunk_ab.py > gpurun_out/k1_chunk_ab_r2a.log 2>&1; echo "chunk_ab exit $?" >> gpurun_out/summary_r2a.txt
cat gpurun_out/summary_r2a.txt; grep -v "Warning\|fork\|^$\|outs =" gpurun_out/pytest_r2a.log | tail -15 | cut -c1-300; cat gpurun_out/k1_chunk_ab_r2a.log | tail -40

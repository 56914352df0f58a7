#!/bin/bash
# round 2, visit n: K1 persistent pass B (conflict-free spline copies) + packed f32x2: parity and A/B
TAG=r2n
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_clahe.py tests/test_gpu_hub.py tests/test_gpu_pipeline.py -s -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python tools/k1_pack_ab.py > gpurun_out/k1_pack_ab_$TAG.log 2>&1; echo "pack ab exit $?" >> gpurun_out/summary_$TAG.txt
cat gpurun_out/summary_$TAG.txt; tail -n 4 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300; cat gpurun_out/k1_pack_ab_$TAG.log

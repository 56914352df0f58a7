#!/bin/bash
# round 2, visit b: sharded-path emulation tests, K3 bound check, wrappers; K3 timing after the seed-striping change
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sharded_emulation.py tests/test_gpu_topk.py tests/test_gpu_hub.py tests/test_gpu_pipeline.py tests/test_gpu_map.py -q -m gpu --timeout 900 -s > gpurun_out/pytest_r2b.log 2>&1; echo "pytest exit $?" > gpurun_out/summary_r2b.txt
timeout 600 python tools/quick_bench.py topk > gpurun_out/quick_topk_r2b.log 2>&1; echo "quick exit $?" >> gpurun_out/summary_r2b.txt
cat gpurun_out/summary_r2b.txt; grep -E "K3 bound check|passed|failed|Error|error" gpurun_out/pytest_r2b.log | tail -30 | cut -c1-250; cat gpurun_out/quick_topk_r2b.log | tail

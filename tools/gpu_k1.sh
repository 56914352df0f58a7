#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/k1_ab.log
timeout 600 python -m pytest tests/test_gpu_clahe.py tests/test_gpu_hub.py -q -m gpu --timeout 300 > gpurun_out/pytest_k1.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 200 python tools/k1_ab.py >> gpurun_out/k1_ab.log 2>&1
timeout 300 python tools/prof_target.py clahe > gpurun_out/prof_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'clahe' -c 4 -o gpurun_out/prof_k1 -f python tools/prof_target.py clahe > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_k1.log; cat gpurun_out/k1_ab.log

#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_gem.py tests/test_gpu_pipeline.py tests/test_gpu_hub.py -q -m gpu --timeout 300 > gpurun_out/pytest_k2.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
GDT_DEBUG_POOL_SINGLE=0 timeout 200 python tools/k2_ab.py > gpurun_out/k2_ab.log 2>&1
timeout 300 python tools/prof_target.py gem > gpurun_out/prof_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_k2.csv python tools/prof_target.py gem > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gem_pool|gem_finalize|whiten' -c 12 -o gpurun_out/prof_k2 -f python tools/prof_target.py gem > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/pytest_k2.log; cat gpurun_out/k2_ab.log

#!/bin/bash
# small-output visit: K1 odd-size timings, launch list of the profiling target, full capture of the K5 kernels only
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/k1_ab.log
timeout 300 python tools/k1_ab.py >> gpurun_out/k1_ab.log 2>&1; echo "k1_ab exit $?" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_clahe.py -q -m gpu --timeout 300 > gpurun_out/pytest_k1.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1q.csv python tools/prof_target.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu_launches exit $?" >> gpurun_out/summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'resize_|reduce_kernel' -c 6 -o gpurun_out/prof_k5_r1q -f python tools/prof_target.py resize > gpurun_out/ncu_full.log 2>&1; echo "ncu_k5 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_k1.log; cat gpurun_out/k1_ab.log; du -sh gpurun_out

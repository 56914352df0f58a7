#!/bin/bash
# round 2, visit k: ncu --set full of K1 pass B variants (non-persistent scalar / persistent / packed) and pass A
TAG=r2k
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
V="0,0,0,0,4,0,0 0,0,0,0,4,1,0 0,0,0,0,4,0,1"
timeout 300 python tools/prof_k1.py $V > gpurun_out/prof_plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'clahe_apply_kernel|clahe_hist_kernel' -s 2 -c 6 -o gpurun_out/prof_k1_$TAG python tools/prof_k1.py $V > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary_$TAG.txt
ls -la gpurun_out/*.ncu-rep; cat gpurun_out/summary_$TAG.txt; tail -3 gpurun_out/ncu_full_$TAG.log

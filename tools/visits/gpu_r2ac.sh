#!/bin/bash
# round 2, visit ac: ncu --set full of K1 v3 (compressed record + persistent pass B), scalar and packed
TAG=r2ac
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
V="0,0,0,-1,4,1,0"
timeout 300 python tools/prof_k1.py $V > gpurun_out/prof_plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'clahe_apply_kernel|clahe_hist_kernel' -s 2 -c 2 -o gpurun_out/prof_k1_$TAG python tools/prof_k1.py $V > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary_$TAG.txt
ls -la gpurun_out/*.ncu-rep; cat gpurun_out/summary_$TAG.txt; tail -3 gpurun_out/ncu_full_$TAG.log

#!/bin/bash
# round 2, visit x: ncu --set full of score_filter_kernel at d = 512 (main pass) -- is the epilogue or the tensor pipe the limiter?
TAG=r2x
mkdir -p gpurun_out
timeout 300 python tools/prof_topk512.py > gpurun_out/prof_plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'score_filter_kernel' -s 2 -c 2 -o gpurun_out/prof_k3_512_$TAG python tools/prof_topk512.py > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_k3_512_$TAG.ncu-rep; tail -2 gpurun_out/ncu_full_$TAG.log

#!/bin/bash
# round 2, visit p: seed-range cap A/B on one N=8-sized shard (125 k x 2048 and 1.25 M x 512)
TAG=r2p
mkdir -p gpurun_out; rm -f gpurun_out/seed_div_$TAG.log
for DIV in 0 64 32 16; do
  for SHAPE in "125000 2048" "1250000 512"; do
    echo "== GDT_DEBUG_K3_SEED_DIV=$DIV shard $SHAPE" >> gpurun_out/seed_div_$TAG.log
    GDT_DEBUG_K3_SEED_DIV=$DIV timeout 300 python tools/search_breakdown.py $SHAPE 10000 2>&1 | grep -E "status|filter|finalize|total" >> gpurun_out/seed_div_$TAG.log
  done
done
cat gpurun_out/seed_div_$TAG.log

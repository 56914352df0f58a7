#!/bin/bash
# round 2, visit r (8-GPU box): NCCL parity tests, contract bench at N=8 (both retrieval configs), per-phase search breakdown
TAG=r2r
N=${1:-8}
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_topk.py tests/test_gpu_sharded_emulation.py tests/test_gpu_map.py -q -m gpu --timeout 300 > gpurun_out/pytest_multi_$TAG.log 2>&1; echo "pytest multi exit $?" >> gpurun_out/summary_$TAG.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err; echo "bench n=$N exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/search_breakdown.py 10000000 512 10000 > gpurun_out/breakdown_10M512_${TAG}_n$N.log 2>&1; echo "breakdown 10M exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tools/search_breakdown.py 1000000 2048 10000 > gpurun_out/breakdown_1M2048_${TAG}_n$N.log 2>&1; echo "breakdown 1M exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_${TAG}_n$N.log 2>&1; echo "h2d ceiling exit $?" >> gpurun_out/summary_$TAG.txt
cat gpurun_out/summary_$TAG.txt; tail -n 12 gpurun_out/h2d_ceiling_${TAG}_n$N.log
tail -n 5 gpurun_out/pytest_multi_$TAG.log | cut -c1-200
cut -c1-400 gpurun_out/scale_${TAG}_n$N.json; tail -n 5 gpurun_out/scale_${TAG}_n$N.err
tail -n 14 gpurun_out/breakdown_10M512_${TAG}_n$N.log; tail -n 14 gpurun_out/breakdown_1M2048_${TAG}_n$N.log

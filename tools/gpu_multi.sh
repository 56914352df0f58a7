#!/bin/bash
# multi-GPU visit: parity suite on GPU 0, then the contract bench under torchrun with N ranks
N=${1:-2}
TAG=${2:-rX}
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest_gpu exit $?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench n=$N exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench n=1 exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "bench ref exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 30 gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/bench_${TAG}_n$N.json; tail -n 8 gpurun_out/bench_${TAG}_n$N.err
cat gpurun_out/bench_${TAG}_n1.json
cat gpurun_out/bench_${TAG}_ref.json

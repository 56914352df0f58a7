#!/bin/bash
# round 2, visit c: re-run of the failed tests + new mining test, per-shard search breakdown, host topology, first bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sharded_emulation.py tests/test_gpu_hub.py tests/test_gpu_pipeline.py tests/test_gpu_map.py tests/test_gpu_multi.py -q -m gpu --timeout 900 -s > gpurun_out/pytest_r2c.log 2>&1; echo "pytest exit $?" > gpurun_out/summary_r2c.txt
for cfg in "125000 2048" "1000000 2048" "1250000 512"; do timeout 300 python tools/search_breakdown.py $cfg >> gpurun_out/search_breakdown_r2c.log 2>&1; done
(nvidia-smi topo -m; echo; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))") > gpurun_out/topology_r2c.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench exit $?" >> gpurun_out/summary_r2c.txt
cat gpurun_out/summary_r2c.txt; grep -E "K3 bound|passed|failed|FAILED" gpurun_out/pytest_r2c.log | tail -12 | cut -c1-200; cat gpurun_out/search_breakdown_r2c.log; cat gpurun_out/topology_r2c.log | tail -25; tail -3 gpurun_out/bench_r2c.err; cat gpurun_out/bench_r2c.json | cut -c1-6000

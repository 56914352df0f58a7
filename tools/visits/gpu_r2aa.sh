#!/bin/bash
# round 2, visit aa: pass A input prefetch (parity + A/B table), K4 timing after the single-search change
TAG=r2ae
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_clahe.py tests/test_gpu_map.py -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python tools/k1_pack_ab.py > gpurun_out/k1_pack_ab_$TAG.log 2>&1; echo "pack ab exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python - > gpurun_out/k4_time_$TAG.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from bench import roxford_shaped
from gandtr_b200.retrieval import PreparedGroundTruth, ShardedIndex, compute_map_and_print
dev = torch.device("cuda", 0)
rq, rdb, rgnd = roxford_shaped()
idx = ShardedIndex(torch.from_numpy(rdb).to(dev)); qd = torch.from_numpy(rq).to(dev)
prep = PreparedGroundTruth("roxford5k", rgnd, idx.n_total, dev)
f = lambda: compute_map_and_print("roxford5k", idx, qd, prep, printer=lambda *_: None)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
print("roxford-shaped mAP E/M/H, prepared ground truth: %.3f ms" % (e0.elapsed_time(e1) / 20))
PY
echo "k4 exit $?" >> gpurun_out/summary_$TAG.txt
cat gpurun_out/summary_$TAG.txt; tail -n 3 gpurun_out/pytest_gpu_$TAG.log | cut -c1-200; grep -E "n=128|verified|div1" gpurun_out/k1_pack_ab_$TAG.log | head -14; cat gpurun_out/k4_time_$TAG.log

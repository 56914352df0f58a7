#!/bin/bash
# round 2, visit t: CUDA-graphed prepared mAP evaluation: parity tests + timing (graph vs eager)
TAG=r2t
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_pipeline.py tests/test_gpu_sharded_emulation.py -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary_$TAG.txt
for NG in "" 1; do
GANDTR_B200_NO_GRAPH=$NG timeout 300 python - >> gpurun_out/k4_time_$TAG.log 2>&1 <<'PY'
import os, sys, torch
sys.path.insert(0, ".")
from bench import roxford_shaped
from gandtr_b200.retrieval import PreparedGroundTruth, ShardedIndex, compute_map_and_print
dev = torch.device("cuda", 0)
rq, rdb, rgnd = roxford_shaped()
idx = ShardedIndex(torch.from_numpy(rdb).to(dev)); qd = torch.from_numpy(rq).to(dev)
prep = PreparedGroundTruth("roxford5k", rgnd, idx.n_total, dev)
res = {}
def f(): res["a"] = compute_map_and_print("roxford5k", idx, qd, prep, printer=lambda *_: None)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
print("roxford-shaped mAP E/M/H, prepared ground truth, NO_GRAPH=%r: %.3f ms  %s  graph_failed=%s" % (os.environ.get("GANDTR_B200_NO_GRAPH"), e0.elapsed_time(e1) / 20, {k: round(float(v), 6) for k, v in res["a"][0].items()}, prep.prepared._graph_failed))
PY
done
echo "k4 exit $?" >> gpurun_out/summary_$TAG.txt
cat gpurun_out/summary_$TAG.txt; tail -n 3 gpurun_out/pytest_gpu_$TAG.log | cut -c1-200; cat gpurun_out/k4_time_$TAG.log

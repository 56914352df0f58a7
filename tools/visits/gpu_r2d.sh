#!/bin/bash
# round 2, visit d: K4 rewrite, mining, store, protocols; roxford-shaped timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_map.py tests/test_gpu_pipeline.py tests/test_gpu_hub.py -q -m gpu --timeout 900 > gpurun_out/pytest_r2d.log 2>&1; echo "pytest exit $?" > gpurun_out/summary_r2d.txt
timeout 600 python - > gpurun_out/k4_r2d.log 2>&1 <<'PY'
import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from bench import roxford_shaped
from gandtr_b200.retrieval import ShardedIndex, compute_map_and_print
from gandtr_b200 import _lib
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for d in (512, 2048):
    rq, rdb, rgnd = roxford_shaped(d=d)
    idx = ShardedIndex(torch.from_numpy(rdb).cuda()); qd = torch.from_numpy(rq).cuda()
    ms = timeit(lambda: compute_map_and_print("roxford5k", idx, qd, rgnd, printer=lambda *_: None))
    probes = torch.randint(0, rdb.shape[0], (70, 110), device="cuda")
    ps = _lib.probe_scores(qd, idx.shard.db, probes)
    out = torch.zeros((70, 110), dtype=torch.int64, device="cuda")
    ms_k = timeit(lambda: _lib.rank_counts(qd, idx.shard.db, probes, ps, out=out))
    print("roxford-shaped d=%d: compute_map_and_print %.3f ms; rank_counts alone (70 x 110 probes x %d rows) %.3f ms" % (d, ms, rdb.shape[0], ms_k))
big = torch.randn((1000000, 2048), device="cuda"); big /= big.norm(dim=1, keepdim=True)
qd = torch.randn((70, 2048), device="cuda"); qd /= qd.norm(dim=1, keepdim=True)
probes = torch.randint(0, 1000000, (70, 110), device="cuda")
ps = _lib.probe_scores(qd, big, probes)
out = torch.zeros((70, 110), dtype=torch.int64, device="cuda")
ms_k = timeit(lambda: _lib.rank_counts(qd, big, probes, ps, out=out), iters=3, warm=1)
print("70 x 1M x 2048: rank_counts %.3f ms (%.0f GB/s of database reads)" % (ms_k, 9 * 8.192e9 / ms_k / 1e6))
PY
echo "k4 exit $?" >> gpurun_out/summary_r2d.txt
cat gpurun_out/summary_r2d.txt; grep -E "passed|failed|FAILED" gpurun_out/pytest_r2d.log | tail -8 | cut -c1-200; cat gpurun_out/k4_r2d.log | tail

#!/bin/bash
# round 2, visit o: early-exit rank merge (parity + timing), launch list of one N=8-sized shard search (125 k x 2048)
TAG=r2o
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_topk.py tests/test_gpu_sharded_emulation.py -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python - > gpurun_out/merge_time_$TAG.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, ".")
from gandtr_b200 import _lib
g, nq, k = 8, 10000, 100
s = torch.randn((g, nq, k), device="cuda").sort(dim=2, descending=True).values
i = (torch.arange(g, device="cuda").view(g, 1, 1) * 125000 + torch.randperm(125000, device="cuda")[:k].sort().values.view(1, 1, k)).expand(g, nq, k).contiguous()
keys = _lib.topk_pack(s, i)
for name, kk in (("sorted lists (rank path)", keys), ("unsorted lists (bitonic path)", keys.flip(2).contiguous())):
    for _ in range(3): _lib.topk_merge_packed(kk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): _lib.topk_merge_packed(kk)
    e1.record(); torch.cuda.synchronize()
    print("merge of %d x %d x %d packed lists, %s: %.3f ms" % (g, nq, k, name, e0.elapsed_time(e1) / 20))
a = _lib.topk_merge_packed(keys); b = _lib.topk_merge_packed(keys.flip(2).contiguous())
print("rank path == bitonic path:", bool(torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])))
PY
echo "merge time exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python tools/search_breakdown.py 125000 2048 10000 > gpurun_out/breakdown_shard_$TAG.log 2>&1; echo "breakdown exit $?" >> gpurun_out/summary_$TAG.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_shard_$TAG.csv python tools/search_breakdown.py 125000 2048 10000 > gpurun_out/ncu_shard_$TAG.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary_$TAG.txt
cat gpurun_out/summary_$TAG.txt; tail -n 3 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300; cat gpurun_out/merge_time_$TAG.log; cat gpurun_out/breakdown_shard_$TAG.log | tail -12

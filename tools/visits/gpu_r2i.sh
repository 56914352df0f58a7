#!/bin/bash
# round 2, visit i: gather-cost microbenchmark, K1 packed-f32x2 A/B + parity, packed-merge timing
TAG=r2i
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
timeout 300 tools/microbench/gather_rate > gpurun_out/gather_rate_$TAG.log 2>&1; echo "gather exit $?" >> gpurun_out/summary_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_clahe.py tests/test_gpu_topk.py tests/test_gpu_sharded_emulation.py tests/test_gpu_pipeline.py -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary_$TAG.txt
timeout 300 python tools/k1_pack_ab.py > gpurun_out/k1_pack_ab_$TAG.log 2>&1; echo "pack ab exit $?" >> gpurun_out/summary_$TAG.txt
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
cat gpurun_out/summary_$TAG.txt; cat gpurun_out/gather_rate_$TAG.log; tail -n 4 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300; cat gpurun_out/k1_pack_ab_$TAG.log gpurun_out/merge_time_$TAG.log

#!/bin/bash
# round 2, visit e: K5 dp4a / batch kernels and K4 fp32-bound path: parity + timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_map.py tests/test_gpu_pipeline.py tests/test_gpu_hub.py -q -m gpu --timeout 900 > gpurun_out/pytest_r2f.log 2>&1; echo "pytest exit $?" > gpurun_out/summary_r2f.txt
timeout 600 python - > gpurun_out/k4k5_r2f.log 2>&1 <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
from bench import roxford_shaped, synth_images_torch
from gandtr_b200.retrieval import ShardedIndex, compute_map_and_print
from gandtr_b200.loader import DeviceImageLoader
from gandtr_b200 import _lib
lib = _lib.load()
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
photos = [synth_images_torch(1, 900 + i, "cuda", h=2304, w=3072)[0] for i in range(8)]
ld = DeviceImageLoader(imsize=1024, device="cuda")
for name, flag in (("dp4a", 0), ("bytewise", 1)):
    lib.gdt_debug_k5_bytewise(flag)
    ms_b = timeit(lambda: ld.resize_batch(photos)) / 8
    ms_1 = timeit(lambda: [ld.resize(p) for p in photos]) / 8
    print("K5 %-8s 3072x2304 -> 1024x768: %.4f ms/image batched (%.0f GB/s), %.4f ms/image one by one" % (name, ms_b, 23.59296 / ms_b, ms_1))
lib.gdt_debug_k5_bytewise(0)
for d in (512, 2048):
    rq, rdb, rgnd = roxford_shaped(d=d)
    idx = ShardedIndex(torch.from_numpy(rdb).cuda()); qd = torch.from_numpy(rq).cuda()
    ms = timeit(lambda: compute_map_and_print("roxford5k", idx, qd, rgnd, printer=lambda *_: None))
    probes = torch.randint(0, rdb.shape[0], (70, 110), device="cuda")
    ps = _lib.probe_scores(qd, idx.shard.db, probes)
    out = torch.zeros((70, 110), dtype=torch.int64, device="cuda")
    ms_k = timeit(lambda: _lib.rank_counts(qd, idx.shard.db, probes, ps, out=out))
    lib.gdt_debug_k4_exact(1)
    ms_x = timeit(lambda: _lib.rank_counts(qd, idx.shard.db, probes, ps, out=out), iters=3, warm=1)
    lib.gdt_debug_k4_exact(0)
    print("roxford-shaped d=%d: compute_map_and_print %.3f ms; rank_counts %.3f ms (every pair exact: %.3f ms)" % (d, ms, ms_k, ms_x))
big = torch.randn((1000000, 2048), device="cuda"); big /= big.norm(dim=1, keepdim=True)
qd = torch.randn((70, 2048), device="cuda"); qd /= qd.norm(dim=1, keepdim=True)
probes = torch.randint(0, 1000000, (70, 110), device="cuda")
ps = _lib.probe_scores(qd, big, probes)
out = torch.zeros((70, 110), dtype=torch.int64, device="cuda")
ms_k = timeit(lambda: _lib.rank_counts(qd, big, probes, ps, out=out), iters=3, warm=1)
print("70 x 1M x 2048: rank_counts %.3f ms (%.0f GB/s of database reads, 9 passes)" % (ms_k, 9 * 8.192e9 / ms_k / 1e6))
PY
echo "k4k5 exit $?" >> gpurun_out/summary_r2f.txt
cat gpurun_out/summary_r2f.txt; grep -E "passed|failed|FAILED" gpurun_out/pytest_r2f.log | tail -8 | cut -c1-200; cat gpurun_out/k4k5_r2f.log | tail -12

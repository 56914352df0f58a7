"""Timing of K1 (run with GDT_DEBUG_K1_OCC=4/6/8 to compare pass-B occupancy variants)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from gandtr_b200 import _lib
from bench import synth_images_torch, MEAN, STD
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for n in (32, 128):
    x = synth_images_torch(n, 1, "cuda")
    out = torch.empty((n, 3, 768, 1024), dtype=torch.float32, device="cuda")
    ms = timeit(lambda: _lib.clahe_u8(x, MEAN, STD, out=out))
    print("occ=%s n=%d: %.3f ms  %.0f img/s  %.0f GB/s algorithmic" % (os.environ.get("GDT_DEBUG_K1_OCC", "6"), n, ms, n / ms * 1e3, n * 15 * 768 * 1024 / ms / 1e6))

import sys, torch
sys.path.insert(0, ".")
from gandtr_b200 import _lib
from bench import MEAN, STD
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
xf = torch.rand((128, 3, 768, 1024), device="cuda") * 2 - 1
half = [0.5, 0.5, 0.5]
out = torch.empty_like(xf)
ms = timeit(lambda: _lib.clahe_f32(xf, half, half, half, half, clip_limit=1.0, out=out))
print("clahe_f32 128 x 1024x768: %.3f ms  %.0f GB/s of 24 B/px  frac %.3f" % (ms, 128*24*768*1024/ms/1e6, 128*24*768*1024/ms/1e6/6554.2))

"""A/B timing of K2 variants (run twice with GDT_DEBUG_POOL_SINGLE=0/1)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from gandtr_b200 import _lib
def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
p = torch.tensor([3.0], device="cuda")
for n, c, h, w in [(128, 2048, 24, 32), (128, 512, 48, 64)]:
    fm = torch.rand((n, c, h, w), device="cuda")
    P = torch.randn((c, c), device="cuda") / c ** 0.5
    m = torch.rand(c, device="cuda") * 0.05
    t_pool = timeit(lambda: _lib.gem_pool(fm, p))
    t_gem = timeit(lambda: _lib.gem_whiten([fm], p, aggregate=True))
    Ps = _lib.whiten_prepare(P)
    t_simt = timeit(lambda: _lib.gem_whiten([fm], p, aggregate=True, P=P, m=m))
    t_all = timeit(lambda: _lib.gem_whiten([fm], p, aggregate=True, P=P, m=m, P_split=Ps))
    gb = fm.numel() * 4 / 1e9
    print("single=%s n=%d c=%d: pool %.1f us (%.0f GB/s)  pool+finalize %.1f us  +whiten(mma.sync) %.1f us  +whiten(tcgen05) %.1f us (%.0f GB/s)" % (
        os.environ.get("GDT_DEBUG_POOL_SINGLE", "0"), n, c, t_pool * 1e3, gb / t_pool * 1e3, t_gem * 1e3, t_simt * 1e3, t_all * 1e3, gb / t_all * 1e3))

# multi-scale (1, 1/sqrt2, 1/2) ResNet-101 shapes
n, c = 128, 2048
fms = [torch.rand((n, c, h, w), device="cuda") for h, w in ((24, 32), (17, 23), (12, 16))]
P = torch.randn((c, c), device="cuda") / c ** 0.5
m = torch.rand(c, device="cuda") * 0.05
Ps = _lib.whiten_prepare(P)
t = timeit(lambda: _lib.gem_whiten(fms, p, aggregate=True, msp_is_p=True, P=P, m=m, P_split=Ps))
gb = sum(f.numel() for f in fms) * 4 / 1e9
print("multi-scale n=%d c=%d: %.1f us (%.0f GB/s)" % (n, c, t * 1e3, gb / t * 1e3))

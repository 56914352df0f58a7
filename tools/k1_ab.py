"""A/B timing of K1's work-split / pipe variants in one process (gdt_debug_k1_config) at 128 images."""
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
lib = _lib.load()
n = 128
x = synth_images_torch(n, 1, "cuda")
out = torch.empty((n, 3, 768, 1024), dtype=torch.float32, device="cuda")
best = None
for chroma_a, texab, occ_a in ((0, 0, 4),):
    for fytex in (0,):
        for spltex in (0,):
            _lib.check(lib.gdt_debug_k1_config(texab, spltex, fytex, chroma_a, occ_a), "cfg")
            ms = timeit(lambda: _lib.clahe_u8(x, MEAN, STD, out=out))
            print("chroma_a=%d texab=%d occ_a=%d fytex=%d spltex=%d n=%d: %.3f ms  %.0f img/s  %.0f GB/s algorithmic" % (chroma_a, texab, occ_a, fytex, spltex, n, ms, n / ms * 1e3, n * 15 * 768 * 1024 / ms / 1e6), flush=True)
            if best is None or ms < best[0]: best = (ms, chroma_a, texab, occ_a, fytex, spltex)
_lib.k1_config_default()
for rows in (32, 48, 24, 16, 12, 8, 0):
    _lib.check(lib.gdt_debug_k1_rows(rows), "rows")
    for nn in (128, 32):
        ms = timeit(lambda: _lib.clahe_u8(x[:nn], MEAN, STD, out=out[:nn]))
        print("rows_per_cta=%d n=%d: %.3f ms  %.0f img/s  %.0f GB/s algorithmic" % (rows, nn, ms, nn / ms * 1e3, nn * 15 * 768 * 1024 / ms / 1e6), flush=True)
for hh, ww in ((683, 1024), (768, 1020), (681, 1023)):      # sizes that take the reflect-padded pass A / non-FAST pass B paths
    xo = synth_images_torch(n, 2, "cuda", h=hh, w=ww)
    oo = torch.empty((n, 3, hh, ww), dtype=torch.float32, device="cuda")
    ms = timeit(lambda: _lib.clahe_u8(xo, MEAN, STD, out=oo))
    print("default config %dx%d n=%d: %.3f ms  %.0f img/s  %.0f GB/s algorithmic" % (ww, hh, n, ms, n / ms * 1e3, n * 15 * hh * ww / ms / 1e6), flush=True)
    del xo, oo
print("best: %.3f ms chroma_a=%d texab=%d occ_a=%d fytex=%d spltex=%d" % best)
_lib.k1_config_default()

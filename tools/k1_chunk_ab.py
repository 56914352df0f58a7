"""A/B timing of K1's launch chunking (gdt_debug_k1_chunk): images per (pass A, pass B) launch pair, i.e. whether a chunk's
5 B/px scratch is still L2-resident when pass B reads it. Prints ms per batch and the algorithmic GB/s."""
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
ref = None
for chunk in (0, -1, 6, 9, 12, 16, 18, 19, 24, 27, 32, 37, 64):
    _lib.check(lib.gdt_debug_k1_chunk(chunk), "chunk")
    for nn in (128, 32):
        ms = timeit(lambda: _lib.clahe_u8(x[:nn], MEAN, STD, out=out[:nn]))
        print("chunk=%d n=%d: %.3f ms  %.0f img/s  %.0f GB/s algorithmic" % (chunk, nn, ms, nn / ms * 1e3, nn * 15 * 768 * 1024 / ms / 1e6), flush=True)
    if ref is None:
        ref = out.clone()
    else:
        assert torch.equal(out, ref), "chunked output differs"
_lib.check(lib.gdt_debug_k1_chunk(0), "chunk")
print("outputs identical for every chunking")

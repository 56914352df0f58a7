"""A/B timing of K1: compressed 32-byte lattice record in pass A (gdt_debug_k1_rec32) x pass B persistent CTAs with conflict-free spline copies (gdt_debug_k1_persist) x packed f32x2
arithmetic (gdt_debug_k1_pack); every combination must be bit-identical. Sizes: the bench shape and two odd ones
(scalar-tail pixels / padded tiles)."""
import ctypes, os, sys
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
for hh, ww, n in ((768, 1024, 128), (768, 1024, 32), (768, 1024, 1), (768, 1020, 64), (681, 1023, 64), (1536, 2048, 16)):
    x = synth_images_torch(n, 1, "cuda", h=hh, w=ww)
    ref = None
    for rec32, persist, pack in ((0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1), (0, 0, 0), (1, 1, 0)):
        _lib.check(lib.gdt_debug_k1_rec32(rec32), "rec32")
        _lib.check(lib.gdt_debug_k1_persist(2 * persist), "persist")
        _lib.check(lib.gdt_debug_k1_pack(pack), "pack")
        out = torch.empty((n, 3, hh, ww), dtype=torch.float32, device="cuda")
        ms = timeit(lambda: _lib.clahe_u8(x, MEAN, STD, out=out))
        same = True if ref is None else bool(torch.equal(ref.view(torch.int32), out.view(torch.int32)))
        if ref is None: ref = out
        print("%dx%d n=%d rec32=%d persist=%d pack=%d: %.3f ms  %.0f img/s  %.0f GB/s algorithmic  identical=%s" % (ww, hh, n, rec32, persist, pack, ms, n / ms * 1e3, n * 15 * hh * ww / ms / 1e6, same), flush=True)
        assert same
    del ref, out, x
_lib.k1_config_default()
print("one-step division verified:", {s: lib.gdt_debug_k1_div1_verified(ctypes.c_float(s)) for s in (0.229, 0.224, 0.225, 0.5)})
x = synth_images_torch(128, 1, "cuda")
out = torch.empty((128, 3, 768, 1024), dtype=torch.float32, device="cuda")
for div1 in (0, 1, 0, 1):
    _lib.check(lib.gdt_debug_k1_div1(div1), "div1")
    ms = timeit(lambda: _lib.clahe_u8(x, MEAN, STD, out=out))
    print("1024x768 n=128 default config, div1=%d: %.3f ms  %.0f GB/s algorithmic" % (div1, ms, 128 * 15 * 768 * 1024 / ms / 1e6))
ref = None
for cf in (0, 1, 0, 1):
    _lib.check(lib.gdt_debug_k1_chroma_f(cf), "chroma_f")
    o2 = torch.empty_like(out)
    ms = timeit(lambda: _lib.clahe_u8(x, MEAN, STD, out=o2))
    same = True if ref is None else bool(torch.equal(ref.view(torch.int32), o2.view(torch.int32)))
    ref = o2 if ref is None else ref
    print("1024x768 n=128 default config, chroma_f=%d: %.3f ms  %.0f GB/s algorithmic  identical=%s" % (cf, ms, 128 * 15 * 768 * 1024 / ms / 1e6, same))
    assert same
_lib.k1_config_default()

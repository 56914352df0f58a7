"""K5 timing: 3072x2304 -> 1024x768 thumbnails in batches of 8, planar horizontal kernel (default) against the interleaved one."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from gandtr_b200 import _lib
from gandtr_b200.loader import DeviceImageLoader
from bench import synth_images_torch
lib = _lib.load()
dev = torch.device("cuda", 0)
photos = [synth_images_torch(1, 900 + i, dev, h=2304, w=3072)[0] for i in range(8)]
ld = DeviceImageLoader(imsize=1024, device=dev)
ref = None
for planar in (0, 1, 0, 1):
    _lib.check(lib.gdt_debug_k5_planar(planar), "planar")
    for _ in range(3): out = ld.resize_batch(photos)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out = ld.resize_batch(photos)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10 / 8
    same = True if ref is None else bool(torch.equal(ref, out))
    ref = out if ref is None else ref
    print("planar=%d: %.1f us per photo  %.0f GB/s algorithmic  frac %.3f  identical=%s" % (planar, ms * 1e3, 23.59296e6 / ms / 1e6, 23.59296e6 / ms / 1e6 / 6554.2, same))
_lib.check(lib.gdt_debug_k5_planar(1), "planar")

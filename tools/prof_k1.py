"""ncu driver for K1 only: one launch pair per variant named on the command line:
texab,spltex,fytex,chroma_a,occ_a[,persist[,pack]] (gdt_debug_k1_config / _persist / _pack)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from gandtr_b200 import _lib
from bench import synth_images_torch, MEAN, STD
lib = _lib.load()
x = synth_images_torch(32, 1, "cuda")
out = torch.empty((32, 3, 768, 1024), dtype=torch.float32, device="cuda")
_lib.clahe_u8(x, MEAN, STD, out=out)
for arg in (sys.argv[1:] or ["0,0,0,0,4"]):
    v = [int(t) for t in arg.split(",")]
    _lib.check(lib.gdt_debug_k1_config(*v[:5]), "cfg")
    _lib.check(lib.gdt_debug_k1_persist(v[5] if len(v) > 5 else 1), "persist")
    _lib.check(lib.gdt_debug_k1_pack(v[6] if len(v) > 6 else 0), "pack")
    _lib.clahe_u8(x, MEAN, STD, out=out)
torch.cuda.synchronize()
_lib.k1_config_default()
lib.gdt_debug_k1_persist(1); lib.gdt_debug_k1_pack(0)
print("done")

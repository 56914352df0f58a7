"""ncu driver for K1 only: one launch pair per pipe variant named on the command line (texab,spltex,fytex,chroma_a,occ_a)."""
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
    _lib.check(lib.gdt_debug_k1_config(*[int(v) for v in arg.split(",")]), "cfg")
    _lib.clahe_u8(x, MEAN, STD, out=out)
torch.cuda.synchronize()
_lib.k1_config_default()
print("done")

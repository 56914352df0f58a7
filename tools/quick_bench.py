"""Ad-hoc kernel timings used while developing (not the contract bench: see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from gandtr_b200 import _lib
from tests.util import MEAN, STD

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

which = sys.argv[1:] or ["clahe", "gem", "topk"]
if "clahe" in which:
    n = 64
    x = torch.randint(0, 256, (n, 768, 1024, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty((n, 3, 768, 1024), dtype=torch.float32, device="cuda")
    ms = timeit(lambda: _lib.clahe_u8(x, MEAN, STD, out=out))
    print("clahe n=%d: %.3f ms  %.0f img/s  %.0f GB/s algorithmic" % (n, ms, n / ms * 1e3, n * 15 * 768 * 1024 / ms / 1e6))
if "gem" in which:
    for (n, c, sizes) in [(64, 512, [(48, 64)]), (64, 2048, [(24, 32)]), (64, 2048, [(24, 32), (17, 23), (12, 16)])]:
        fm = [torch.rand((n, c, h, w), device="cuda") for h, w in sizes]
        p = torch.tensor([3.0], device="cuda")
        P = torch.randn((c, c), device="cuda") / c ** 0.5
        m = torch.rand(c, device="cuda") * 0.05
        agg = len(sizes) > 1
        ms = timeit(lambda: _lib.gem_whiten(fm, p, aggregate=True, msp_is_p=agg, P=P, m=m))
        byts = sum(f.numel() * 4 for f in fm)
        print("gem+whiten n=%d c=%d scales=%d: %.3f ms %.0f img/s %.0f GB/s" % (n, c, len(sizes), ms, n / ms * 1e3, byts / ms / 1e6))
        p2 = torch.tensor([2.92], device="cuda")
        ms = timeit(lambda: _lib.gem_whiten(fm, p2, aggregate=True, msp_is_p=agg, P=P, m=m))
        print("   p=2.92: %.3f ms %.0f GB/s" % (ms, byts / ms / 1e6))
if "topk" in which:
    for (nq, ndb, d) in [(1024, 131072, 512), (10000, 125000, 2048), (10000, 1000000, 512), (10000, 1000000, 2048)]:
        db = torch.randn((ndb, d), device="cuda"); db /= db.norm(dim=1, keepdim=True)
        q = torch.randn((nq, d), device="cuda"); q /= q.norm(dim=1, keepdim=True)
        shadow, nmax = _lib.db_prepare(db)
        ws = torch.empty(_lib.score_topk_workspace_bytes(nq, ndb, d, 100), dtype=torch.uint8, device="cuda")
        res = [None]
        def f(): res[0] = _lib.score_topk(q, db, shadow, nmax, 100, ws=ws)
        ms = timeit(f, iters=5, warm=2)
        st = res[0][2].cpu().numpy()
        print("topk nq=%d ndb=%d d=%d: %.3f ms  %.1f TFLOP/s  %.0f q/s status=%s" % (nq, ndb, d, ms, 2.0 * nq * ndb * d / ms / 1e9, nq / ms * 1e3, st))
        del db, q, shadow, ws

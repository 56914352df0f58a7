"""Short driver for ncu captures: a few launches of every hot kernel at bench shapes (smaller batch)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from gandtr_b200 import _lib
from bench import synth_images_torch, MEAN, STD

which = sys.argv[1:] or ["clahe", "resize", "gem", "topk", "map", "chain"]
dev = torch.device("cuda", 0)
reps = 3
if "clahe" in which:
    x = synth_images_torch(32, 1, dev)
    out = torch.empty((32, 3, 768, 1024), dtype=torch.float32, device=dev)
    for _ in range(reps):
        _lib.clahe_u8(x, MEAN, STD, out=out)
if "resize" in which:
    from gandtr_b200.loader import DeviceImageLoader
    ph = synth_images_torch(1, 900, dev, h=2304, w=3072)[0]
    ld = DeviceImageLoader(imsize=1024, device=dev)
    for _ in range(reps):
        ld.resize(ph)
    photos = [synth_images_torch(1, 910 + i, dev, h=2304, w=3072)[0] for i in range(8)]
    for _ in range(reps):
        ld.resize_batch(photos)                                   # one launch per pass for 8 images
    big = synth_images_torch(1, 901, dev, h=3456, w=4608)[0]      # pre-reduction by 2, then LANCZOS
    ld.resize(big)
if "gem" in which:
    c = 2048
    fm = [torch.rand((128, c, h, w), device=dev) for h, w in ((24, 32), (17, 23), (12, 16))]
    p = torch.tensor([3.0], device=dev)
    P = torch.randn((c, c), device=dev) / c ** 0.5
    m = torch.rand(c, device=dev) * 0.05
    Ps = _lib.whiten_prepare(P)
    for _ in range(reps):
        _lib.gem_whiten(fm[:1], p, aggregate=True, P=P, m=m, P_split=Ps)
    _lib.gem_whiten(fm, p, aggregate=True, msp_is_p=True, P=P, m=m)
    fv = torch.rand((32, 512, 48, 64), device=dev)
    _lib.gem_whiten([fv], torch.tensor([2.92], device=dev), aggregate=True)
if "topk" in which:
    nq, ndb, d = 4096, 262144, 2048
    db = torch.randn((ndb, d), device=dev); db /= db.norm(dim=1, keepdim=True)
    q = torch.randn((nq, d), device=dev); q /= q.norm(dim=1, keepdim=True)
    shadow, nmax = _lib.db_prepare(db)
    for _ in range(reps):
        s, i, st = _lib.score_topk(q, db, shadow, nmax, 100)
    print("status", st.cpu().tolist())
if "map" in which:
    from bench import roxford_shaped
    from gandtr_b200.retrieval import PreparedGroundTruth, ShardedIndex, compute_map_and_print
    rq, rdb, rgnd = roxford_shaped()
    idx = ShardedIndex(torch.from_numpy(rdb).to(dev))
    qd = torch.from_numpy(rq).to(dev)
    prep = PreparedGroundTruth("roxford5k", rgnd, idx.n_total, dev)
    for _ in range(reps):
        compute_map_and_print("roxford5k", idx, qd, prep, printer=lambda *_: None)
if "chain" in which:
    xf = torch.rand((32, 3, 768, 1024), device=dev) * 2 - 1
    half = [0.5, 0.5, 0.5]
    for _ in range(reps):
        y = _lib.clahe_f32(xf, half, half, half, half, clip_limit=1.0)
        _lib.meanstd_adapt(y, half, half, MEAN, STD, out=y)
torch.cuda.synchronize()
print("done")

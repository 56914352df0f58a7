"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference) in the build container.
The reference cannot travel to the GPU box, so its outputs on seeded synthetic inputs are committed as small
fixtures together with this script.  Run:  PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def synth_image(seed, h, w, kind):
    """The synthetic image families of SURVEY.md 8(d): uniform noise / smooth sinusoid + noise / dark."""
    rs = np.random.RandomState(seed)
    if kind == "noise":
        return rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([128 + 70 * (np.sin(xx / (37.0 + 5 * c)) + np.cos(yy / (23.0 + 3 * c))) for c in range(3)], -1)
    img = img + rs.normal(0, 8, img.shape)
    if kind == "dark":
        img = 255.0 * (np.clip(img, 0, 255) / 255.0) ** 2.8
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    import torch
    from PIL import Image
    hub = ref_harness.load_reference()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    ref = hub.gem_vgg16_hedngan(pretrained=False, device="cpu")

    # ---- K1: full transform, bit-exact target -------------------------------------------------------
    cases = [("noise", 64, 96, 11), ("smooth", 96, 128, 12), ("smooth", 61, 77, 13), ("dark", 100, 130, 14),
             ("smooth", 72, 100, 15), ("noise", 37, 53, 16)]
    pack = {}
    for i, (kind, h, w, seed) in enumerate(cases):
        img = synth_image(seed, h, w, kind)
        out = ref.transform(Image.fromarray(img)).numpy()
        pack["img%d" % i] = img
        pack["out%d" % i] = out
    np.savez_compressed(os.path.join(GOLD, "clahe_transform.npz"), **pack)

    # ---- ClahePost (float input on a normalised CHW tensor) -----------------------------------------
    from mdir.components.data import wrapper as W
    post = W.ClahePost("[[0.5,0.5,0.5],[0.5,0.5,0.5]]", "1.0", device="cpu")
    rs = np.random.RandomState(21)
    x = (rs.rand(3, 48, 80).astype(np.float32) * 2.2 - 1.1)  # slightly outside [-1,1]: exercises cv2's clipping
    y = post.postprocess(torch.from_numpy(x.copy()), None, None).numpy()
    post4 = W.ClahePost("[[0.485,0.456,0.406],[0.229,0.224,0.225]]", "4.0", device="cpu")
    x2 = ((synth_image(22, 45, 67, "smooth").astype(np.float32) / 255.0).transpose(2, 0, 1)
          - np.array([0.485, 0.456, 0.406], np.float32)[:, None, None]) / np.array([0.229, 0.224, 0.225], np.float32)[:, None, None]
    x2 = x2 + rs.normal(0, 0.01, x2.shape).astype(np.float32)
    y2 = post4.postprocess(torch.from_numpy(x2.copy()), None, None).numpy()
    np.savez_compressed(os.path.join(GOLD, "clahe_post.npz"), x0=x, y0=y, x1=x2, y1=y2)

    # ---- K2: GeM + L2N, multi-scale aggregation, whitening ---------------------------------------------
    from cirtorch.layers.pooling import GeM
    from cirtorch.layers.normalization import L2N
    pack = {}
    rs = np.random.RandomState(31)
    for ci, (c, sizes, p) in enumerate([(32, [(12, 16), (8, 11), (6, 8)], 3.0), (48, [(7, 9), (5, 6), (3, 4)], 2.92),
                                        (64, [(24, 32)], 3.0)]):
        n = 2
        pool, norm = GeM(p=p), L2N()
        fm = [np.abs(rs.normal(0, 1, (n, c, h, w))).astype(np.float32) * (rs.rand(n, c, 1, 1) > 0.2) for (h, w) in sizes]
        fm = [f.astype(np.float32) for f in fm]
        with torch.no_grad():
            per_scale = [norm(pool(torch.from_numpy(f))).squeeze(-1).squeeze(-1).permute(1, 0) for f in fm]  # D x N each
        P = (rs.normal(0, 1, (c, c)) / np.sqrt(c)).astype(np.float64)
        m = (0.05 * rs.rand(c, 1)).astype(np.float64)
        wh = W.CirtorchWhiten.__new__(W.CirtorchWhiten)
        wh.P = torch.tensor(P, dtype=torch.float32)
        wh.m = torch.tensor(m, dtype=torch.float32)
        wh.dimensions = c - 8
        outs_plain, outs_agg, outs_wh = [], [], []
        for i in range(n):
            cols = [ps[:, i:i + 1].clone() for ps in per_scale]
            outs_plain.append(cols[0].squeeze().numpy())
            agg = W.CirMultiscaleAggregation.aggregate_tensor(cols, len(cols), c, p)
            outs_agg.append(agg.numpy().copy())
            outs_wh.append(wh.postprocess(agg.clone(), None, None).numpy())
        for si, f in enumerate(fm):
            pack["c%d_fmap%d" % (ci, si)] = f
        pack["c%d_p" % ci] = np.float32(p)
        pack["c%d_P" % ci] = P
        pack["c%d_m" % ci] = m
        pack["c%d_dim" % ci] = np.int64(c - 8)
        pack["c%d_plain" % ci] = np.stack(outs_plain)
        pack["c%d_agg" % ci] = np.stack(outs_agg)
        pack["c%d_whiten" % ci] = np.stack(outs_wh)
    np.savez_compressed(os.path.join(GOLD, "descriptors.npz"), **pack)

    # ---- K3/K4: ranking + mAP on planted ground truth ---------------------------------------------------
    from cirtorch.utils.evaluate import compute_map_and_print, compute_map
    rs = np.random.RandomState(41)
    d, ndb, nq = 64, 1500, 12
    db = rs.normal(0, 1, (ndb, d)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    gnd, q = [], []
    for i in range(nq):
        ids = rs.permutation(ndb)[:40]
        ne, nh, nj = rs.randint(3, 12), (0 if i % 5 == 4 else rs.randint(3, 12)), rs.randint(2, 10)
        easy, hard, junk = ids[:ne], ids[ne:ne + nh], ids[ne + nh:ne + nh + nj]
        centre = db[np.concatenate([easy, hard])].mean(0)
        v = centre + 0.35 * rs.normal(0, 1, d) / np.sqrt(d)
        q.append((v / np.linalg.norm(v)).astype(np.float32))
        gnd.append({"bbx": None, "easy": easy, "hard": hard, "junk": junk})
    q = np.stack(q)
    scores = np.dot(db, q.T)                       # cirscore.py:71 layout: ndb x nq
    ranks = np.argsort(-scores, axis=0)            # cirscore.py:72
    avg, per = compute_map_and_print("roxford5k", ranks, gnd)
    gnd_old = [{"ok": np.concatenate([g["easy"], g["hard"]]), "junk": g["junk"]} for g in gnd]
    gnd_old[3]["ok"] = np.array([], dtype=np.int64)  # query without positives -> NaN and excluded
    avg_old, per_old = compute_map_and_print("tokyo", ranks, gnd_old)
    mapM, apsM, mprM, prsM = compute_map(ranks, [{"ok": np.concatenate([g["easy"], g["hard"]]), "junk": g["junk"]} for g in gnd], [1, 5, 10])
    np.savez_compressed(
        os.path.join(GOLD, "map_eval.npz"), db=db, q=q, ranks=ranks.astype(np.int64),
        easy=np.array([np.pad(g["easy"], (0, 40 - len(g["easy"])), constant_values=-1) for g in gnd]),
        hard=np.array([np.pad(g["hard"], (0, 40 - len(g["hard"])), constant_values=-1) for g in gnd]),
        junk=np.array([np.pad(g["junk"], (0, 40 - len(g["junk"])), constant_values=-1) for g in gnd]),
        map_easy=avg["map_easy"], map_medium=avg["map_medium"], map_hard=avg["map_hard"],
        ap_easy=per["ap_easy"], ap_medium=per["ap_medium"], ap_hard=per["ap_hard"],
        old_map=avg_old["map"], old_ap=per_old["ap"], mprM=mprM, prsM=prsM, mapM=mapM)

    # ---- whitening learning -----------------------------------------------------------------------------
    from cirtorch.utils.whiten import whitenlearn, whitenapply
    rs = np.random.RandomState(51)
    D, n = 24, 400
    X = rs.normal(0, 1, (D, n))
    X /= np.linalg.norm(X, axis=0, keepdims=True)
    qidxs = list(range(0, 100))
    pidxs = list(range(100, 200))
    X[:, pidxs] = X[:, qidxs] + 0.3 * rs.normal(0, 1, (D, 100)) / np.sqrt(D)
    m, P = whitenlearn(X, qidxs, pidxs)
    Y = whitenapply(X[:, :50], m, P, dimensions=16)
    np.savez_compressed(os.path.join(GOLD, "whiten.npz"), X=X, qidxs=np.array(qidxs), pidxs=np.array(pidxs), m=m, P=P, Y=Y)

    # ---- end-to-end hub model on a tiny image (random-init weights, shared through state_dict) ----------
    img = synth_image(61, 64, 80, "smooth")
    x = ref.transform(Image.fromarray(img)).unsqueeze(0)
    with torch.no_grad():
        dsc = ref(x).numpy()
    np.savez_compressed(os.path.join(GOLD, "hub_vgg16_tiny.npz"), img=img, desc=dsc,
                        pool_p=ref.model.pool.p.detach().numpy())
    print("golden fixtures written to", os.path.abspath(GOLD))
    for f in sorted(os.listdir(GOLD)):
        print("  %-28s %8d B" % (f, os.path.getsize(os.path.join(GOLD, f))))


if __name__ == "__main__":
    main()

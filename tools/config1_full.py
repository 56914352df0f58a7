"""BASELINE config 1 at its stated size, end to end on one GPU:
    gem_vgg16_hedngan(pretrained=False): CLAHE + extract + whiten + rank, 64 queries vs 1 k synthetic 1024x768 images.
Images: 25 % uniform noise, 75 % smooth sinusoid + N(0, 8) (seeds 1000 + i); query j = database image j with a +-8
brightness jitter, so ground truth is meaningful. Whitening learnt on the 64 (query, source) pairs + the 1 k database
descriptors. Checks, against the oracle (CPU): K1 output bit-exact on three full-size images; descriptors of the same
images against the reference's CPU path (torch CPU fp32, same weights); top-100 index lists identical to the exact
ranking of the GPU's own descriptors; mAP equal. Prints one JSON line (commit it under profiles/).
    python tools/config1_full.py [n_db] [n_q]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from gandtr_b200 import hub, whiten as W
from gandtr_b200.extract import extract_descriptors
from gandtr_b200.retrieval import ShardedIndex, compute_map_and_print
from oracle import clahe_np as O
from oracle import retrieval_np as R
from tests.util import MEAN, STD, load_lut, synth_image

H, Wd = 768, 1024


def db_image(i):
    return synth_image(1000 + i, H, Wd, "noise" if i % 4 == 0 else "smooth")


def main():
    ndb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = hub.gem_vgg16_hedngan(pretrained=False)
    dev = torch.device("cuda", 0)
    t0 = time.time()
    images = [db_image(i) for i in range(ndb)]
    qimages = [np.clip(images[j].astype(np.int16) + (8 if j % 2 else -8), 0, 255).astype(np.uint8) for j in range(nq)]
    gnd = [{"ok": np.array([j]), "junk": np.array([], dtype=np.int64)} for j in range(nq)]
    t_synth = time.time() - t0

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        return r, e0.elapsed_time(e1)

    extract_descriptors(net, images[:8], None, net.transform, batch_size=8)         # warm-up (cuDNN autotune, gdt_init)
    dbv, ms_db = timed(lambda: extract_descriptors(net, images, None, net.transform, batch_size=8))
    qv, ms_q = timed(lambda: extract_descriptors(net, qimages, None, net.transform, batch_size=8))

    # whitening on the (query, source) pairs (float64, reference's whitenlearn contract), then projection to 512-d
    X = torch.cat([dbv, qv]).double().t().contiguous()                                # D x (ndb + nq)
    qidxs, pidxs = list(range(ndb, ndb + nq)), list(range(nq))
    (m, P), ms_learn = timed(lambda: W.whitenlearn(X, qidxs, pidxs))
    Y = W.whitenapply(X, m, P).t().contiguous().float()                               # (ndb + nq) x D, L2-normalised
    dbw, qw = Y[:ndb].contiguous().to(dev), Y[ndb:].contiguous().to(dev)

    index = ShardedIndex(dbw)
    k = min(100, ndb)
    (s, i), ms_search = timed(lambda: index.search(qw, k))
    (avg, aps), ms_map = timed(lambda: compute_map_and_print("config1", index, qw, gnd, printer=lambda *_: None))

    # ---- parity against the oracle ----
    lut = load_lut()
    spot = [0, 1, 2]                                                                  # one noise image, two smooth ones
    k1_bad = 0
    for j in spot:
        got = net.transform.batch(torch.from_numpy(images[j][None]).to(dev))[0].cpu().numpy()
        ref = O.transform_u8(images[j], lut, MEAN, STD)
        k1_bad += int((got.view(np.uint32) != ref.view(np.uint32)).sum())
    # reference CPU path for the same three images: the same modules on the CPU in fp32 (stock conv, GeM, L2N)
    from oracle import descriptors_np as D
    import copy
    inner = getattr(net, "model", net)
    feats = copy.deepcopy(inner.features).cpu().float()
    p = float(inner.pool.p.detach().cpu().reshape(-1)[0])
    desc_diff = 0.0
    with torch.no_grad():
        for j in spot:
            x = torch.from_numpy(O.transform_u8(images[j], lut, MEAN, STD))[None]
            fm = feats(x).numpy().astype(np.float64)
            ref = D.forward_descriptor(fm, p).reshape(-1)
            desc_diff = max(desc_diff, float(np.abs(dbv[j].double().cpu().numpy() - ref).max() / np.abs(ref).max()))
    sc = R.scores_exact(qw.cpu().numpy(), dbw.cpu().numpy())
    os_, oi = R.topk(sc, k)
    lists_equal = bool(np.array_equal(oi, i.cpu().numpy()))
    ranks = np.argsort(-sc.astype(np.float64), axis=1, kind="stable").T
    omap = R.compute_map(ranks, gnd)[0]
    out = {
        "config": "BASELINE config 1: gem_vgg16_hedngan(pretrained=False), %d queries vs %d synthetic 1024x768 images" % (nq, ndb),
        "extract_db_ms": ms_db, "extract_images_per_s": ndb / ms_db * 1e3, "extract_q_ms": ms_q,
        "whitenlearn_ms": ms_learn, "search_top%d_ms" % k: ms_search, "map_ms": ms_map,
        "map": float(avg["map"]), "oracle_map": float(omap), "map_equal": bool(abs(float(avg["map"]) - float(omap)) < 1e-9),
        "k1_mismatching_floats_on_3_images": k1_bad,
        "descriptor_max_rel_diff_vs_cpu_fp32_backbone": desc_diff,
        "top%d_lists_identical_to_exact_ranking" % k: lists_equal,
        "host_image_synthesis_s": t_synth,
    }
    print(json.dumps(out))
    assert k1_bad == 0 and lists_equal and out["map_equal"]


if __name__ == "__main__":
    main()

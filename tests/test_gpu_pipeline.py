"""Rows a16-a22 of SURVEY.md 8(a) on the GPU: batched extraction, whitening learning, the CirDatasetAp score object and
the validate / infer stage functions, on a small synthetic dataset (random-init VGG16, BASELINE config 1 in miniature)."""
import numpy as np
import pytest
import torch

from oracle import retrieval_np as R
from oracle import whiten_np as WO
from tests.util import golden, synth_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vgg():
    from gandtr_b200 import hub
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    return hub.gem_vgg16_hedngan(pretrained=False)


def _dataset(ndb=24, nq=5):
    sizes = [(64, 80), (64, 80), (72, 64), (64, 80)]
    images = [synth_image(1000 + i, *sizes[i % len(sizes)], "smooth" if i % 4 else "noise") for i in range(ndb)]
    qimages, gnd = [], []
    for j in range(nq):
        img = np.clip(images[j].astype(np.int16) + (8 if j % 2 else -8), 0, 255).astype(np.uint8)   # brightness jitter
        qimages.append(img)
        gnd.append({"ok": np.array([j]), "junk": np.array([(j + 7) % ndb])})
    return images, qimages, gnd


def test_batched_extraction_equals_per_image_forward(vgg):
    from gandtr_b200.extract import extract_descriptors, extract_vectors
    images, _, _ = _dataset()
    d = extract_descriptors(vgg, images, None, vgg.transform, batch_size=4)
    assert tuple(d.shape) == (len(images), 512) and d.is_cuda
    with torch.no_grad():
        for i in (0, 2, 5, 23):
            ref = vgg(vgg.transform(images[i]).unsqueeze(0)).reshape(-1)
            torch.testing.assert_close(d[i], ref, rtol=2e-5, atol=5e-6)   # cuDNN picks other conv algorithms for other batch sizes
    v = extract_vectors(vgg, images[:6], None, vgg.transform, print_freq=0)
    assert tuple(v.shape) == (512, 6) and not v.is_cuda                   # reference convention: D x n on the host
    torch.testing.assert_close(v.t(), d[:6].cpu(), rtol=2e-5, atol=5e-6)
    # sharded extraction: the two halves are exactly the rows of the full matrix
    halves = [extract_descriptors(vgg, images, None, vgg.transform, rank=r, world_size=2) for r in range(2)]
    torch.testing.assert_close(torch.cat(halves), d, rtol=2e-5, atol=5e-6)   # batch boundaries move -> other cuDNN algorithms


def test_extraction_is_deterministic_across_repeats_and_batchings(vgg):
    """Race detector for the double-buffered upload / K1 / K2 chain: the same images give the same bits, run after run, and
    rows do not depend on which other images shared their launch (uint8 -> CLAHE is per image; cuDNN picks algorithms per
    batch size, hence the tolerance on the second half)."""
    from gandtr_b200.extract import extract_descriptors
    images, qimages, _ = _dataset()
    first = extract_descriptors(vgg, images, None, vgg.transform)
    for _ in range(4):
        assert torch.equal(extract_descriptors(vgg, images, None, vgg.transform), first)
    q = extract_descriptors(vgg, qimages, None, vgg.transform)
    for _ in range(4):
        assert torch.equal(extract_descriptors(vgg, qimages, None, vgg.transform), q)
    torch.testing.assert_close(extract_descriptors(vgg, images, None, vgg.transform, batch_size=1), first, rtol=2e-5, atol=5e-6)
    x = vgg.transform.batch(torch.from_numpy(np.stack([images[0], images[1]])).cuda())
    for _ in range(4):
        assert torch.equal(vgg.transform.batch(torch.from_numpy(np.stack([images[0], images[1]])).cuda()), x)


def test_device_resize_extraction_equals_host_resize(vgg):
    """with `device_resize` the workers only decode; crop + LANCZOS thumbnail run on the GPU (K5) and feed K1
    the very same pixels, so the descriptors are those of the host-resize path."""
    from gandtr_b200.extract import extract_descriptors, load_image
    from gandtr_b200.loader import DeviceImageLoader
    images = [synth_image(1200 + i, *[(150, 200), (200, 150), (96, 128)][i % 3], "smooth" if i % 3 else "noise") for i in range(9)]
    bbxs = [None, (10, 20, 140, 190), None, None, (5.5, 4.5, 120.2, 190.7), None, None, None, (0, 0, 100, 90)]
    ld = DeviceImageLoader(imsize=96, device="cuda")
    from PIL import Image
    pil = [Image.fromarray(a) for a in images]
    for img, bbx in zip(pil, bbxs):
        assert np.array_equal(ld.load(img, bbx=bbx).cpu().numpy(), load_image(img.copy(), 96, bbx))
        assert ld.load(img, bbx=bbx).shape[0] <= 96
    host = extract_descriptors(vgg, pil, 96, vgg.transform, bbxs=bbxs, batch_size=4, device_resize=False)
    dev = extract_descriptors(vgg, pil, 96, vgg.transform, bbxs=bbxs, batch_size=4)       # default: K5 does the geometry
    assert torch.equal(host, dev), "max |diff| %g in rows %s" % (float((host - dev).abs().max()),
                                                                  (host != dev).any(dim=1).nonzero().flatten().tolist())
    # arrays pass through unresized on both paths (datahelpers.py:76-79)
    assert torch.equal(extract_descriptors(vgg, images[:3], 96, vgg.transform, device_resize=False),
                       extract_descriptors(vgg, images[:3], 96, vgg.transform, device_resize=True))


def test_extract_ms_matches_reference_formula(vgg):
    from gandtr_b200.extract import extract_ms
    img = synth_image(5, 96, 128, "smooth")
    x = vgg.transform(img).unsqueeze(0)
    ms, msp = [1, 1 / np.sqrt(2), 0.5], 3.0
    with torch.no_grad():
        got = extract_ms(vgg, x, ms, msp)
        v = torch.zeros(512, dtype=torch.float64)
        for s in ms:
            xs = x if s == 1 else torch.nn.functional.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False)
            v += vgg.model(xs).double().cpu().reshape(-1).pow(msp)          # imageretrievalnet.py:348-353
        v = (v / len(ms)).pow(1.0 / msp)
        v /= v.norm()
    np.testing.assert_allclose(got.numpy(), v.numpy(), rtol=2e-5, atol=2e-7)


def test_whitenlearn_matches_reference_golden_up_to_sign():
    from gandtr_b200 import whiten as W
    g = golden("whiten.npz")
    m, P = W.whitenlearn(g["X"], g["qidxs"], g["pidxs"])
    m, P = m.cpu().numpy(), P.cpu().numpy()
    np.testing.assert_allclose(m, g["m"], rtol=1e-12, atol=1e-14)
    sign = np.sign((P * g["P"]).sum(1, keepdims=True))
    np.testing.assert_allclose(P * sign, g["P"], rtol=1e-6, atol=1e-8)
    Y = W.whitenapply(g["X"][:, :50], m, P, dimensions=16).cpu().numpy()
    np.testing.assert_allclose(Y * sign[:16], g["Y"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(Y.T @ Y, g["Y"].T @ g["Y"], rtol=1e-6, atol=1e-9)     # every inner product is unchanged
    # stage function with named vectors (stages/whiten.py:30-75)
    names = ["v%d" % i for i in range(g["X"].shape[1])]
    meta, whit = W.learn_lw_whitening({}, (names, g["X"].T.copy(), [names[i] for i in g["qidxs"]], [names[i] for i in g["pidxs"]]))
    assert meta["stats"]["failed_times"] == 0 and whit["P"].shape == g["P"].shape
    mo, Po = WO.whitenlearn(g["X"], g["qidxs"], g["pidxs"])
    np.testing.assert_allclose(np.abs(whit["P"]), np.abs(Po), rtol=1e-6, atol=1e-8)


def test_cirdatasetap_and_validate_stage(vgg):
    from gandtr_b200.extract import extract_descriptors
    from gandtr_b200.score import SCORES, validate
    images, qimages, gnd = _dataset()
    ds = {"name": "synthetic", "images": images, "qimages": qimages, "bbxs": [None] * len(qimages), "gnd": gnd}
    data = vgg.network_params.runtime["data"]
    score = SCORES["cirdatasetap"]({"image_size": None, "dataset": ds, "transforms": data.get("transforms", data.get("augmentations")),
                                    "mean_std": data["mean_std"]})
    logged = []
    avg = score(vgg, "cuda", lambda it, size, label, value, dtype: logged.append((it, size, label, value, dtype)))
    # oracle: descriptors -> exact scores -> full ranking -> compute_map
    db = extract_descriptors(vgg, images, None, vgg.transform).cpu().numpy()
    q = extract_descriptors(vgg, qimages, None, vgg.transform).cpu().numpy()
    m, aps, _, _ = R.compute_map(R.full_ranks(R.scores_exact(q, db)), gnd)
    assert abs(avg["map"] - m) < 1e-12
    assert avg["map"] > 0.9                                     # brightness-jittered copies must be found
    labels = [x[2] for x in logged]
    assert labels[0] == "dataset" and labels[1] == "score_avg" and labels.count("score") == len(qimages)
    assert logged[1][4] == "scalar/score" and score.decisive_criterion == "val/learning/score_avg:map_medium"
    out, = validate({"network": vgg, "data": {},
                     "validation": {"type": "MultiCriterialValidation", "decisive_criterion": None,
                                    "synthetic": {"type": "SingleValidation", "frequency": None, "network_overlay": None,
                                                  "data": None, "criterion": {"type": "cirdatasetap", "image_size": None,
                                                                              "dataset": ds}}}})
    assert abs(out["eval"]["synthetic/validation/score_avg:map"] - m) < 1e-12


def test_cirdatasetap_runs_the_networks_eval_wrappers(vgg):
    """`extract_vectors(network, ...)` calls `net(input)` (imageretrievalnet.py:326-343), so a SingleNetwork's eval
    wrappers -- {0_cirwhiten, 1_cirmultiscale} for the pretrained hub models (embedding.yml:24-25) -- shape the
    descriptors CirDatasetAp ranks. Batched extraction must equal the per-image `net(transform(img))` descriptors and
    validate() and infer() must agree on them."""
    from gandtr_b200 import network as N
    from gandtr_b200.extract import extract_descriptors
    from gandtr_b200.score import SCORES, infer
    rs = np.random.RandomState(11)
    c = 512
    whit = {"P": rs.normal(0, 1, (c, c)) / np.sqrt(c), "m": 0.05 * rs.rand(c, 1)}
    data = vgg.network_params.runtime["data"]
    runtime = {"data": data, "wrappers": {"train": None, "eval": {"0_cirwhiten": {"whitening": whit, "dimensions": 96},
                                                                  "1_cirmultiscale": {"scales": True}}}}
    net = N.SingleNetwork(vgg.model, N.SingleNetwork.NetworkParams(vgg.network_params.model, runtime), "cuda", frozen=True)
    net.transform = vgg.transform
    images, qimages, gnd = _dataset(12, 4)
    d = extract_descriptors(net, images, None, vgg.transform, batch_size=4)
    assert tuple(d.shape) == (12, 96)                              # whitened, reduced dimensionality
    with torch.no_grad():
        per_image = torch.stack([net(vgg.transform(im).unsqueeze(0)).reshape(-1) for im in images])
        plain = extract_descriptors(vgg, images, None, vgg.transform, batch_size=4)
    torch.testing.assert_close(d, per_image, rtol=1e-4, atol=1e-5)   # batch size changes the cuDNN algorithm
    assert plain.shape[1] == 512                                    # the wrapper-less network is untouched
    # the score object ranks the wrapped descriptors
    ds = {"name": "synthetic", "images": images, "qimages": qimages, "bbxs": [None] * len(qimages), "gnd": gnd}
    score = SCORES["cirdatasetap"]({"image_size": None, "dataset": ds, "transforms": data.get("transforms", data.get("augmentations")),
                                    "mean_std": data["mean_std"]})
    avg = score(net, "cuda", lambda *a: None)
    q = extract_descriptors(net, qimages, None, vgg.transform).cpu().numpy()
    m, _, _, _ = R.compute_map(R.full_ranks(R.scores_exact(q, d.cpu().numpy())), gnd)
    assert abs(avg["map"] - m) < 1e-12
    # infer() goes through the same wrappers: both stages see the same vectors
    _, _, vecs = infer({"network": net}, (images[:3],))
    np.testing.assert_allclose(vecs, d[:3].double().cpu().numpy(), rtol=1e-4, atol=1e-5)
    # the reference rejects unknown criterion keys (`assert not params`, cirscore.py:48)
    with pytest.raises(AssertionError):
        SCORES["cirdatasetap"]({"image_size": None, "dataset": ds, "transforms": data.get("transforms", data.get("augmentations")),
                                "mean_std": data["mean_std"], "multiscale": True})


def test_infer_stage_and_whitening_learning_chain(vgg, tmp_path):
    from gandtr_b200.score import infer, infer_and_learn_whitening
    images, _, _ = _dataset(12, 2)
    meta, names, vecs = infer({"network": vgg}, (images,))
    assert vecs.dtype == np.float64 and vecs.shape == (12, 512) and len(names) == 12
    np.testing.assert_allclose(np.linalg.norm(vecs, axis=1), 1.0, atol=1e-5)
    cids = ["%06d" % i for i in range(12)]
    pkl = {"cids": cids, "images": images, "qidxs": [0, 1, 2, 3], "pidxs": [4, 5, 6, 7]}
    meta, whit = infer_and_learn_whitening({"network": vgg, "whitening": {"type": "pca", "dataset_pkl": pkl, "directory": str(tmp_path)}})
    assert whit["P"].shape == (512, 512) and whit["m"].shape == (512, 1)
    assert meta["whitening_path"].endswith("whitening/pca-memory.pkl")
    meta2, whit2 = infer_and_learn_whitening({"network": vgg, "whitening": {"type": "pca", "dataset_pkl": pkl, "directory": str(tmp_path)}})
    assert meta2["status"] == "skipped" and whit2 is None


def test_hard_negative_mining_equals_reference_walk():
    """the K3-based search against the reference algorithm (full sort + per-query cluster walk,
    traindataset.py:246-279) restated with the oracle's total order."""
    from gandtr_b200.mining import search_hard_negatives
    from tests.util import unit_rows
    rs = np.random.RandomState(4)
    d, nq, npool, nnum = 128, 40, 6000, 5
    pool = unit_rows(rs, npool, d)
    q = unit_rows(rs, nq, d)
    poolclusters = rs.randint(0, 300, npool)
    poolclusters[:200] = 7                                  # a big cluster: forces deep walks for cluster-7 neighbours
    qclusters = rs.randint(0, 300, nq)
    q[:5] = pool[:5] + 0.05 * rs.normal(0, 1, (5, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    nidxs, stats = search_hard_negatives(torch.from_numpy(q.T.copy()).cuda(), torch.from_numpy(pool.T.copy()).cuda(),
                                         qclusters, poolclusters, nnum, depth=8)
    ranks = R.full_ranks(R.scores_exact(q, pool))           # [npool, nq]
    for qi in range(nq):
        clusters, ref = {qclusters[qi]}, []
        r = 0
        while len(ref) < nnum:
            cand = ranks[r, qi]
            if poolclusters[cand] not in clusters:
                ref.append(int(cand))
                clusters.add(poolclusters[cand])
            r += 1
        assert nidxs[qi] == ref
    assert len(stats["average_negative_distance"]) == nq * nnum


def test_file_formats_checkpoint_whitening_gnd_and_jpeg_paths(vgg, tmp_path, monkeypatch):
    """SURVEY 8(b) 'File formats': a reference-layout checkpoint + whitening pickle load through the pretrained hub path,
    gnd_<dataset>.pkl drives CirDatasetAp, and images come from JPEG files on disk (PIL decode, LANCZOS thumbnail)."""
    import pickle
    from PIL import Image
    from gandtr_b200 import hub
    from gandtr_b200.score import CirDatasetAp
    # --- checkpoint in the reference layout (mdir/learning/network.py:212-220) + whitening pickle (stages/whiten.py:75)
    wdir = tmp_path / "weights"
    wdir.mkdir()
    torch.save(vgg.state_dict(), str(wdir / "hedngan_embed_vgg16.pth"))
    rs = np.random.RandomState(3)
    whit = {"m": 0.05 * rs.rand(512, 1), "P": rs.normal(0, 1, (512, 512)) / np.sqrt(512)}
    with open(str(wdir / "hedngan_embed_vgg16_lw.pkl"), "wb") as f:
        pickle.dump(whit, f)
    net = hub.gem_vgg16_hedngan(pretrained=True, weights_dir=str(wdir))
    assert [type(w).__name__ for w in net.wrappers["eval"].wrappers] == ["CirtorchWhiten", "CirMultiscaleAggregation"]
    for k, v in vgg.model.state_dict().items():
        assert torch.equal(v, net.model.state_dict()[k])
    img = synth_image(3, 96, 128, "smooth")
    with torch.no_grad():
        vec = net(net.transform(img).unsqueeze(0))
    assert tuple(vec.shape) == (512,) and abs(float(vec.norm()) - 1.0) < 1e-5      # whitened multi-scale descriptor
    # --- dataset on disk: JPEG files + gnd pickle (cirtorch/datasets/testdataset.py:6-38)
    root = tmp_path / "data" / "test" / "roxford5k"
    (root / "jpg").mkdir(parents=True)
    names = ["im%03d" % i for i in range(10)]
    for i, n in enumerate(names):
        Image.fromarray(synth_image(50 + i, 120, 160, "smooth")).save(str(root / "jpg" / (n + ".jpg")), quality=95)
    gnd = [{"bbx": [10, 10, 150, 110], "easy": [0], "hard": [5], "junk": [9]}, {"bbx": None, "easy": [1], "hard": [], "junk": []}]
    with open(str(root / "gnd_roxford5k.pkl"), "wb") as f:
        pickle.dump({"imlist": names, "qimlist": names[:2], "gnd": gnd}, f)
    monkeypatch.setenv("GANDTR_DATA_ROOT", str(tmp_path / "data"))
    data = vgg.network_params.runtime["data"]
    score = CirDatasetAp({"image_size": 128, "dataset": "roxford5k", "transforms": data.get("transforms", data.get("augmentations")),
                          "mean_std": data["mean_std"]})
    assert score.bbxs == [(10, 10, 150, 110), None] and len(score.images) == 10
    avg = score(vgg, "cuda", lambda *a: None)
    assert set(avg) == {"map_easy", "map_medium", "map_hard"}
    # oracle on the same descriptors (queries are cropped / resized by the same loader)
    from gandtr_b200.extract import extract_descriptors
    db = extract_descriptors(vgg, score.images, 128, vgg.transform).cpu().numpy()
    q = extract_descriptors(vgg, score.qimages, 128, vgg.transform, bbxs=score.bbxs).cpu().numpy()
    oavg, _, _ = R.compute_map_protocols("roxford5k", R.full_ranks(R.scores_exact(q, db)), gnd)
    for k in avg:
        assert abs(avg[k] - oavg[k]) < 1e-12


def test_diverse_anchor_mining_equals_reference_loop():
    """SURVEY 8(f) N3, second half: the device loop (gdt_diverse_anchors) against the reference algorithm
    (cirtorch_datasets.py:78-96: matrix-vector product, running maximum, argsort slice, random choice) restated on the
    host with the library's exact scores and total order, driven by the same random draws."""
    from gandtr_b200.mining import diverse_anchor_ranks, mark_easy_pairs, select_diverse_anchors
    from tests.util import unit_rows
    rs = np.random.RandomState(6)
    d, npool, qsize = 96, 1500, 120
    centres = unit_rows(rs, 40, d)
    pool = centres[rs.randint(0, 40, npool)] + 0.35 * rs.normal(0, 1, (npool, d)).astype(np.float32) / np.sqrt(d)
    pool[700] = pool[3]                                               # exact duplicates: ties broken by index
    pool = (pool / np.linalg.norm(pool, axis=1, keepdims=True)).astype(np.float32)
    qvecs = torch.from_numpy(pool.T.copy()).cuda()
    for shuffle, (excl, incl) in ((True, (0.02, 0.3)), (False, (0.0, 0.1)), (True, (0.0, 1.0))):
        gen = torch.Generator().manual_seed(123)
        idxs, scores = select_diverse_anchors(qvecs, qsize, excl, incl, shuffle=shuffle, generator=gen)
        ranks = diverse_anchor_ranks(npool, qsize, excl, incl, shuffle, torch.Generator().manual_seed(123))
        # reference loop
        S = R.scores_exact(pool, pool)                               # [npool, npool], fp64-accumulated, rounded once
        idx, ref_idxs, ref_scores = 0, [0], []
        most = None
        for t in range(qsize - 1):
            dist = S[:, idx]
            most = dist.copy() if most is None else np.maximum(most, dist)
            order = np.lexsort((np.arange(npool), most))             # value asc, index asc
            idx = int(order[ranks[t]])
            ref_scores.append(float(most[idx]))
            ref_idxs.append(idx)
        assert idxs == ref_idxs
        assert scores == ref_scores
        assert len(set(idxs)) >= qsize - 1                           # only the planted exact duplicate can repeat (as in the reference)
    easy = mark_easy_pairs(qvecs[:, :20], qvecs[:, 20:40], 0.25)
    sim = (pool[:20] * pool[20:40]).sum(1)
    assert easy.count("-easy") == 5 and all((lab == "-easy") == (i in set(np.argsort(sim)[-5:].tolist())) for i, lab in enumerate(easy))


def test_descriptor_store_roundtrip_into_sharded_index(tmp_path):
    """SURVEY 8(f) N4 on the device: descriptors written by 3 'ranks', read back as 2 shards through the pinned
    double-buffered loader, searched through ShardedIndex -- identical to searching the matrix that was saved; the bf16
    payload reproduces the bf16-rounded rows exactly and still ranks the planted neighbours first."""
    from gandtr_b200 import store
    from gandtr_b200.retrieval import ShardedIndex, shard_bounds
    from tests.util import run_ranks, unit_rows
    rs = np.random.RandomState(10)
    n, d, nq, k = 50000, 128, 40, 20
    db = unit_rows(rs, n, d)
    src = rs.randint(0, n, nq)
    q = db[src] + 0.3 * rs.normal(0, 1, (nq, d)).astype(np.float32) / np.sqrt(d)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    dbd = torch.from_numpy(db).cuda()
    for dtype in ("float32", "bfloat16"):
        path = str(tmp_path / dtype)
        for r in range(3):
            lo, hi = shard_bounds(n, 3, r)
            store.save(path, dbd[lo:hi], ids=["img%d" % i for i in range(n)] if r == 0 else None, rows_per_file=7000,
                       lo=lo, n_total=n, dtype=dtype)
        man = store.load_manifest(path)
        assert man["rows"] == n and man["dtype"] == dtype
        parts = [store.load_shard(path, r, 2, device="cuda", block_rows=3000) for r in range(2)]
        full = torch.cat([p[0] for p in parts])
        assert parts[1][1] == shard_bounds(n, 2, 1)[0] and parts[0][2] == n and full.is_cuda
        if dtype == "float32":
            assert torch.equal(full, dbd)
        else:
            assert torch.equal(full, dbd.to(torch.bfloat16).to(torch.float32))
            assert float((full - dbd).abs().max()) < 2e-3
        expect = full.cpu().numpy()
        os_, oi = R.topk(R.scores_exact(q, expect), k)

        def rank_fn(rank, comm):
            rows, lo, total = store.load_shard(path, rank, 2, device="cuda")
            index = ShardedIndex(rows, n_total=total, index_base=lo, comm=comm)
            s, i = index.search(torch.from_numpy(q).cuda(), k)
            assert np.array_equal(i.cpu().numpy(), oi) and np.abs(s.cpu().numpy() - os_).max() < 1e-6
            return int((i[:, 0].cpu() == torch.from_numpy(src)).sum())
        hits = run_ranks(2, rank_fn)
        assert hits[0] == nq                                          # planted neighbours rank first, also from bf16 rows


@pytest.mark.parametrize("d,n", [(64, 1000), (70, 333), (512, 30011), (130, 33), (2048, 4100), (1, 5)])
def test_syrk_f64_gram_products(d, n):
    """the library's own fp64 rank-n update behind whitenlearn / pcawhitenlearn against the float64
    product (the reference's np.dot(df, df.T), whiten.py:20,42,47): symmetric, bit-reproducible, strided input."""
    from gandtr_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(d * 7 + n)
    A = torch.randn((d, n), generator=g, device="cuda", dtype=torch.float64)
    C = _lib.syrk_f64(A, 0.5)
    ref = 0.5 * (A @ A.t())
    torch.testing.assert_close(C, ref, rtol=1e-12, atol=1e-12)
    assert torch.equal(C, C.t())
    assert torch.equal(C, _lib.syrk_f64(A, 0.5))                      # fixed summation order: same bits every run
    wide = torch.randn((d, n + 9), generator=g, device="cuda", dtype=torch.float64)
    view = wide[:, 4:4 + n]                                           # row stride != n
    torch.testing.assert_close(_lib.syrk_f64(view), view @ view.t(), rtol=1e-12, atol=1e-12)

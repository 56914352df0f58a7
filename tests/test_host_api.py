"""CPU-side checks of the host mirror: registries, string grammar, reprs, C-ABI symbol table, and that the product
fails loudly instead of falling back when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def test_library_exports_every_declared_symbol():
    from gandtr_b200 import _lib
    header = open(os.path.join(ROOT, "include", "gandtr_b200.h")).read()
    declared = set(re.findall(r"\b(gdt_[a-z0-9_]+)\s*\(", header))
    lib = _lib.load()
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gdt_abi_version() == 1
    assert lib.gdt_status_string(-7) == b"score_topk candidate overflow"


def test_transform_registry_grammar_and_repr():
    from gandtr_b200 import transforms as T
    t = T.initialize_transforms("pil2np | apply_clahe:1.0 | totensor | normalize", [[0.485, 0.456, 0.406], [0.229, 0.224, 0.225]])
    assert [type(x).__name__ for x in t.transforms] == ["Pil2Numpy", "ApplyClahe", "ToTensor", "Normalize"]
    assert t.transforms[1].params == {"clip_limit": 1.0, "grid_size": 8, "colorspace": "lab"}
    assert repr(t.transforms[1]) == "ApplyClahe(clip_limit=1.0, grid_size=8, colorspace=lab)"
    t2 = T.initialize_transforms("pil2np|apply_clahe:2.5:4:lab|totensor|normalize", [[0.5] * 3, [0.5] * 3])
    assert t2.transforms[1].params["clip_limit"] == 2.5 and t2.transforms[1].params["grid_size"] == 4
    with pytest.raises(NotImplementedError):
        T.initialize_transforms("pil2np | mirror | totensor | normalize", [[0.5] * 3, [0.5] * 3])
    with pytest.raises(NotImplementedError):
        T.initialize_transforms("pil2np | apply_clahe:1.0:8:luv | totensor | normalize", [[0.5] * 3, [0.5] * 3])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from gandtr_b200 import _lib, hub, transforms as T
    t = T.initialize_transforms("pil2np | apply_clahe:1.0 | totensor | normalize", [[0.5] * 3, [0.5] * 3])
    with pytest.raises(_lib.GdtError):
        t(np.zeros((16, 16, 3), np.uint8))
    with pytest.raises(RuntimeError):
        hub.gem_vgg16_hedngan(pretrained=False)
    with pytest.raises(RuntimeError):
        hub.gem_vgg16_hedngan(pretrained=False, device="cpu")
    with pytest.raises(_lib.GdtError):
        _lib.gem_whiten([torch.zeros(1, 4, 2, 2)], torch.ones(1))
    lib = _lib.load()
    buf = (ctypes.c_int16 * (33 * 33 * 33 * 3))()
    assert lib.gdt_init(buf) == -3                                                # GDT_ERR_NO_DEVICE


def test_wrapper_registry_and_scales():
    from gandtr_b200 import network as N
    c = N.initialize_wrappers("cirfaketuplebatch", "cpu")
    assert [type(w).__name__ for w in c.wrappers] == ["CirFakeTupleBatch"]
    ms = N.CirMultiscaleAggregation(True, device="cpu")
    np.testing.assert_allclose(ms.scales, [1, 1 / np.sqrt(2), 0.5])
    assert N.CirMultiscaleAggregation("ss", device="cpu").scales == [1]
    assert len(N.CirMultiscaleAggregation("sms5", device="cpu").scales) == 5
    whit = {"P": np.eye(8), "m": np.zeros((8, 1))}
    c2 = N.initialize_wrappers({"1_cirmultiscale": {"scales": True}, "0_cirwhiten": {"whitening": whit, "dimensions": None}}, "cpu")
    assert [type(w).__name__ for w in c2.wrappers] == ["CirtorchWhiten", "CirMultiscaleAggregation"]
    assert c2.wrappers[0].dimensions == 8 and c2.wrappers[0].P.dtype == torch.float32
    with pytest.raises(NotImplementedError):
        N.initialize_wrappers("reflectpad_divisible:32", "cpu")
    # the `augment` wrapper string of finetune.yml:13: commas inside the bracketed arguments do not split (utils.splitp)
    c3 = N.initialize_wrappers("meanstd_post:[[0.5,0.5,0.5],[0.5,0.5,0.5]]:[[0.485,0.456,0.406],[0.229,0.224,0.225]],"
                               "clahepost:[[0.5,0.5,0.5],[0.5,0.5,0.5]]:1.0,cir_ratio_pass_through:0.25:anc", "cpu")
    assert [type(w).__name__ for w in c3.wrappers] == ["MeanStdPost", "ClahePost", "CirRatioPassThrough"]
    assert c3.wrappers[1].clip_limit == 1.0 and c3.wrappers[2].probability == 0.25
    assert repr(c3.wrappers[2]).startswith("CirRatioPassThrough(probability=0.25, train_label=")
    with pytest.raises(AssertionError):
        N.initialize_wrappers("meanstd_post:[[0.5,0.5", "cpu")
    # FakeBatch stacks per-image vectors into D x n (wrapper.py:266-280)
    out = N.FakeBatch("cpu").postprocess([torch.arange(4.0).reshape(4, 1), torch.ones(4)], None, None)
    assert tuple(out.shape) == (4, 2)
    flat, meta = N.CirFakeTupleBatch("cpu").preprocess([[1, 2], [3, 4]], None)
    assert flat == [1, 2, 3, 4] and meta == 2


def test_network_construction_matches_reference_layout():
    from gandtr_b200 import network as N
    net = N.init_cirnet(local_whitening=False, pooling="gem", regional=False, whitening=False, pretrained=False,
                        cir_architecture="vgg16")
    keys = list(net.state_dict().keys())
    assert keys[0] == "features.0.weight" and "pool.p" in keys and len(keys) == 27
    assert net.meta["outputdim"] == 512 and net.meta["in_channels"] == 3
    assert "meta" in repr(net) and "GeM(p=3.0000, eps=1e-06)" in repr(net)
    with pytest.raises(NotImplementedError):
        N.init_cirnet(local_whitening=False, pooling="gem", regional=False, whitening=False, pretrained=True,
                      cir_architecture="vgg16")
    from gandtr_b200.generator import ResnetGenerator
    g = ResnetGenerator()
    assert sum(p.numel() for p in g.parameters()) == 11378179           # 9-block ResNet generator, instance norm
    assert list(g.state_dict().keys())[0] == "model.1.weight"


def test_descriptor_store_roundtrip_and_resharding(tmp_path):
    """blocks written by 3 'ranks', read back as 2 shards -- rows, ids and whitening survive."""
    from gandtr_b200 import store
    from gandtr_b200.retrieval import shard_bounds
    rs = np.random.RandomState(0)
    n, d = 1000, 16
    x = rs.normal(0, 1, (n, d)).astype(np.float32)
    whit = {"m": rs.rand(d, 1), "P": rs.normal(0, 1, (d, d))}
    for r in range(3):
        lo, hi = shard_bounds(n, 3, r)
        store.save(str(tmp_path), x[lo:hi], ids=["img%d" % i for i in range(n)] if r == 0 else None,
                   whitening=whit if r == 0 else None, rows_per_file=128, lo=lo, n_total=n)
    man = store.load_manifest(str(tmp_path))
    assert man["rows"] == n and man["dim"] == d and man["shards"][0]["lo"] == 0 and man["shards"][-1]["hi"] == n
    parts = []
    for r in range(2):
        rows, lo, total = store.load_shard(str(tmp_path), r, 2, device="cpu")
        assert lo == shard_bounds(n, 2, r)[0] and total == n
        parts.append(rows.numpy())
    np.testing.assert_array_equal(np.concatenate(parts), x)
    w = store.load_whitening(str(tmp_path))
    np.testing.assert_array_equal(w["P"], whit["P"])
    # bf16 payload: half the bytes, rows come back as the bf16-rounded values
    bdir = str(tmp_path / "bf16")
    store.save(bdir, x, rows_per_file=300, dtype="bfloat16")
    assert store.load_manifest(bdir)["dtype"] == "bfloat16"
    rows, lo, total = store.load_shard(bdir, 0, 1, device="cpu")
    np.testing.assert_array_equal(rows.numpy(), torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy())
    with pytest.raises(ValueError):
        store.save(bdir, x, dtype="float16")


def test_metadata_tensor_and_pass_through_routing():
    """tools/tensors.py:8-30,37-85 and wrapper.py:97-146 on the host: structure-preserving helpers, metadata propagation
    and the md5-based routing of CirRatioPassThrough (no kernels involved)."""
    import hashlib
    from gandtr_b200 import network as N
    t = N.MetadataTensor(torch.zeros(1, 3, 4, 4), {"image_label": ["anc"], "name": ["img_000"]})
    assert tuple(t.shape) == (1, 3, 4, 4)                                  # tensor attributes shine through
    moved = t.to("cpu")
    assert isinstance(moved, N.MetadataTensor) and moved.metadata is t.metadata
    up = torch.nn.functional.interpolate(t, scale_factor=2.0)              # __torch_function__ keeps the metadata
    assert isinstance(up, N.MetadataTensor) and tuple(up.tensor.shape) == (1, 3, 8, 8)
    nested = {"a": [t, None], "b": (t, torch.ones(2))}
    bare = N.as_tensor(nested)
    assert isinstance(bare["a"][0], torch.Tensor) and bare["a"][1] is None and isinstance(bare["b"], tuple)
    dev = N.to_device(nested, "cpu")
    assert isinstance(dev["a"][0], N.MetadataTensor) and dev["a"][1] is None
    m2 = N.as_metadata_tensor(torch.ones(2), {"k": 1})
    assert m2.metadata == {"k": 1} and N.as_metadata_tensor(m2, {"j": 2}).metadata == {"k": 1, "j": 2}
    ms = N.CirMultiscaleAggregation(True, device="cpu")
    scaled, waslist = ms.preprocess(N.MetadataTensor(torch.zeros(1, 3, 8, 8), {"name": "x"}), None)
    assert not waslist and all(isinstance(s, N.MetadataTensor) for s in scaled) and tuple(scaled[2].tensor.shape) == (1, 3, 4, 4)
    # routing
    w = N.CirRatioPassThrough("0.25", "anc", device="cpu")
    names = ["img_%03d" % i for i in range(64)]
    expect = [int(hashlib.md5(n.encode("utf8")).hexdigest()[-4:], 16) / 65536 < 0.25 for n in names]
    assert [w._passthrough(n) for n in names] == expect and 4 < sum(expect) < 28
    items = [N.MetadataTensor(torch.full((1,), float(i)), {"image_label": "anc" if i % 2 == 0 else "pos", "name": n})
             for i, n in enumerate(names[:8])]
    through, skipped = w.preprocess(items, None)
    for i in range(8):
        goes = (i % 2 == 0) and expect[i]
        assert (through[i] is not None) == goes and (skipped[i] is None) == goes
    out = w.postprocess([None if x is None else x.tensor + 100 for x in through], None, skipped)
    for i in range(8):
        assert float(out[i]) == (i + 100 if (i % 2 == 0 and expect[i]) else i) and isinstance(out[i], torch.Tensor)
    rp = N.RandomPassThrough("1.0", "cpu")
    assert rp.preprocess(torch.ones(1), None)[1] is None and repr(rp) == "RandomPassThrough(probability=1.0)"


def test_print_scores_mirrors_the_scenario_step(capsys):
    """mdir/examples/perform_scenario.py:19-41: labels, rounding and the ({},) return of the reference's print_scores."""
    from gandtr_b200.score import print_scores
    meta = {"eval": {"roxford5k/validation/score_avg:map_medium": 0.647453, "247tokyo1k/validation/score_avg:map": 0.9,
                     "roxford5k/validation/score_avg:map_easy": 0.7, "val/validation/loss_avg:dist": np.float32(0.125)}}
    assert print_scores({"metadata": meta}, ()) == ({},)
    out = capsys.readouterr().out
    assert "\nEval\n" in out
    assert "    %-20s %s" % ("roxford.5k medium", 64.75) in out and "    %-20s %s" % ("247tokyo.1k", 90.0) in out
    assert "    %-20s %s" % ("dist", 0.125) in out and "map_easy" not in out

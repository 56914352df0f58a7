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

"""Pins oracle/reference_cpu.py (the cv2/torch/NumPy restatement used as CPU baseline) to the golden fixtures made
by the unmodified reference and to the NumPy oracle."""
import numpy as np
import pytest

from oracle import clahe_np, descriptors_np, reference_cpu
from tests.util import MEAN, STD, golden, load_lut

cv2 = pytest.importorskip("cv2")


def test_transform_cv2_bit_exact_vs_reference_goldens():
    g = golden("clahe_transform.npz")
    n = len([k for k in g.files if k.startswith("img")])
    for i in range(n):
        out = reference_cpu.transform_cv2(g["img%d" % i], MEAN, STD)
        assert np.array_equal(out.view(np.uint32), g["out%d" % i].view(np.uint32))


def test_transform_cv2_equals_numpy_oracle():
    rs = np.random.RandomState(5)
    img = rs.randint(0, 256, (80, 104, 3)).astype(np.uint8)
    a = reference_cpu.transform_cv2(img, MEAN, STD)
    b = clahe_np.transform_u8(img, load_lut(), MEAN, STD)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_descriptor_chain_vs_goldens():
    import torch
    g = golden("descriptors.npz")
    for ci in range(3):
        fm = [g[k] for k in sorted(k for k in g.files if k.startswith("c%d_fmap" % ci))]
        p = float(g["c%d_p" % ci])
        per = [reference_cpu.gem_l2n_torch(torch.from_numpy(f), p) for f in fm]
        assert np.allclose(per[0].numpy(), g["c%d_plain" % ci], rtol=1e-6, atol=1e-7)
        agg = reference_cpu.aggregate_torch(per, p)
        assert np.allclose(agg.numpy(), g["c%d_agg" % ci], rtol=1e-5, atol=1e-7)
        wh = reference_cpu.whiten_torch(agg, torch.tensor(g["c%d_P" % ci], dtype=torch.float32),
                                        torch.tensor(g["c%d_m" % ci], dtype=torch.float32), int(g["c%d_dim" % ci]))
        assert np.allclose(wh.numpy(), g["c%d_whiten" % ci], rtol=1e-5, atol=1e-6)
        ref = descriptors_np.descriptor_pipeline(fm, p=p, aggregate=True, msp_is_p=True, P=g["c%d_P" % ci],
                                                 m=g["c%d_m" % ci], dimensions=int(g["c%d_dim" % ci]))
        assert np.allclose(wh.numpy(), ref, rtol=2e-5, atol=2e-6)

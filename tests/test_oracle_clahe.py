"""Pins oracle/clahe_np.py: (1) against golden outputs of the unmodified reference transform, (2) against live
cv2 when the pinned version is installed. Bit-exact."""
import numpy as np
import pytest

from oracle import clahe_np as O
from tests.util import MEAN, STD, golden, load_lut, synth_image


def test_transform_matches_reference_golden_bit_exact():
    g = golden("clahe_transform.npz")
    lut = load_lut()
    n = len([k for k in g.files if k.startswith("img")])
    assert n >= 6
    for i in range(n):
        out = O.transform_u8(g["img%d" % i], lut, MEAN, STD)
        ref = g["out%d" % i]
        assert out.dtype == np.float32 and out.shape == ref.shape
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), "case %d: %d mismatches" % (i, (out != ref).sum())


def test_clahe_post_matches_reference_golden_bit_exact():
    g = golden("clahe_post.npz")
    lut = load_lut()
    y0 = O.clahe_post_f32(g["x0"], lut, [[0.5] * 3, [0.5] * 3], clip_limit=1.0)
    assert np.array_equal(y0.view(np.uint32), g["y0"].view(np.uint32))
    y1 = O.clahe_post_f32(g["x1"], lut, [MEAN, STD], clip_limit=4.0)
    assert np.array_equal(y1.view(np.uint32), g["y1"].view(np.uint32))


def test_shipped_lut_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    if cv2.__version__ != "4.13.0":
        pytest.skip("arithmetic characterised on opencv-python 4.13.0 only")
    assert np.array_equal(O.probe_rgb2lab_lut_cv2(), load_lut())


@pytest.mark.parametrize("kind,h,w", [("noise", 48, 64), ("smooth", 50, 67), ("dark", 40, 56), ("smooth", 33, 31)])
def test_stages_match_live_cv2(kind, h, w):
    cv2 = pytest.importorskip("cv2")
    if cv2.__version__ != "4.13.0":
        pytest.skip("arithmetic characterised on opencv-python 4.13.0 only")
    lut = load_lut()
    x = synth_image(5, h, w, kind).astype(np.float32) / np.float32(255.0)
    lab = cv2.cvtColor(x, cv2.COLOR_RGB2LAB)
    assert np.array_equal(O.rgb2lab_f32(x, lut), lab)
    L8 = ((lab[..., 0] / np.float32(100)) * np.float32(255)).astype(np.uint8)
    for clip in (1.0, 4.0):
        ref = cv2.createCLAHE(clipLimit=clip, tileGridSize=(8, 8)).apply(L8)
        assert np.array_equal(O.clahe_u8(L8, clip, 8), ref)
    assert np.array_equal(O.lab2rgb_f32(lab), cv2.cvtColor(lab, cv2.COLOR_LAB2RGB))


def test_spline_table_is_monotone_and_anchored():
    tab = O.inv_gamma_spline_tab()
    assert tab.shape == (1024, 4) and tab.dtype == np.float32
    assert tab[0, 0] == 0.0 and np.all(np.diff(tab[:, 0]) > 0)

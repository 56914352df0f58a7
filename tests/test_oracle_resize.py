"""Pins oracle/resize_np.py (the restatement of Pillow's thumbnail / reduce / LANCZOS resample reached from
genericdataset.py:86-97 and datahelpers.py:75-82): (1) against fixtures produced by Pillow through the reference's own
call sequence, (2) against live Pillow, (3) the library's host-side geometry / coefficient code against the oracle."""
import ctypes

import numpy as np
import pytest

from oracle import resize_np as R
from tests.util import golden, synth_image


def _cases():
    g = golden("resize.npz")
    n = len([k for k in g.files if k.startswith("img")])
    assert n >= 9
    for i in range(n):
        bbx = tuple(float(v) for v in g["bbx%d" % i]) or None
        yield g["img%d" % i], int(g["imsize%d" % i]), bbx, g["out%d" % i]


def test_oracle_matches_pillow_golden_bit_exact():
    for img, imsize, bbx, ref in _cases():
        out = R.load_resized_u8(img, imsize, bbx)
        assert out.shape == ref.shape and np.array_equal(out, ref), (img.shape, imsize, bbx)


def test_oracle_matches_live_pillow():
    PIL = pytest.importorskip("PIL")
    from PIL import Image
    lanczos = getattr(Image, "LANCZOS", Image.Resampling.LANCZOS)
    rs = np.random.RandomState(3)
    for h, w, s in [(60, 80, 40), (33, 31, 32), (480, 640, 512), (50, 70, 80), (400, 90, 33), (777, 333, 50), (90, 1100, 60),
                    (512, 512, 511), (3, 700, 64), (700, 2, 64)]:
        a = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        im = Image.fromarray(a)
        im.thumbnail((s, s), lanczos)
        assert np.array_equal(np.asarray(im), R.thumbnail_u8(a, s)), (h, w, s)
    a = rs.randint(0, 256, (64, 95, 3)).astype(np.uint8)
    for fx, fy in [(1, 2), (2, 1), (2, 2), (3, 3), (4, 4), (5, 5), (2, 3), (7, 6)]:
        assert np.array_equal(np.asarray(Image.fromarray(a).reduce((fx, fy))), R.reduce_u8(a, fx, fy)), (fx, fy)


def test_library_host_geometry_matches_oracle():
    """gdt_thumbnail_geometry / gdt_debug_resize_coeffs are host code: checked here without a GPU."""
    from gandtr_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(4)
    for _ in range(3000):
        w, h = int(rs.randint(1, 5000)), int(rs.randint(1, 5000))
        imsize = float(rs.choice([rs.randint(1, 2048), rs.uniform(1, 2048)]))
        ow, oh, fx, fy, resized = _lib.thumbnail_geometry(w, h, imsize)
        size = R.thumbnail_size(w, h, imsize)
        if size is None or size == (w, h):
            assert not resized and (ow, oh) == (w, h)
        else:
            assert resized and (ow, oh) == size and (fx, fy) == R.reduce_factors(w, h, *size), (w, h, imsize)
    for in_size, in0, in1, out_size in [(640, 0.0, 640.0, 512), (129, 0.0, 128.75, 64), (1024, 0.0, 1024.0, 724), (83, 0.0, 83.0, 50),
                                        (50, 0.0, 50.0, 50), (7, 0.0, 7.0, 3), (3000, 0.0, 3000.0, 1024)]:
        ksize, bounds, kk = R.precompute_coeffs(in_size, in0, in1, out_size)
        ks = ctypes.c_int()
        b = np.zeros((out_size, 2), np.int32)
        k = np.zeros(out_size * (ksize + 8), np.int32)
        _lib.check(lib.gdt_debug_resize_coeffs(in_size, in0, in1, out_size, ctypes.byref(ks), b.ctypes.data_as(ctypes.c_void_p),
                                               k.ctypes.data_as(ctypes.c_void_p), k.size), "gdt_debug_resize_coeffs")
        assert ks.value == ksize and np.array_equal(b, bounds)
        assert np.array_equal(k[:out_size * ksize].reshape(out_size, ksize), kk)

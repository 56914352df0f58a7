"""The per-pixel arithmetic the K1 kernels inline (gandtr_b200/csrc/clahe_math.cuh), compiled for the host by g++ and
compared bit for bit with oracle/clahe_np.py (itself pinned to the reference's golden outputs). Runs without a GPU:
pass A (uint8 / float pixel -> CLAHE input byte + Q14 chroma) and pass B after the blend (byte + chroma -> normalised
RGB, table and recomputed variants, SIMD-body and scalar-tail sequences)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import clahe_np as O
from tests.util import MEAN, STD, load_lut

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "k1_math_host.cpp")
f32 = np.float32


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("k1h") / "libk1h.so")
    subprocess.run([gxx, "-O2", "-shared", "-fPIC", "-mfma", "-ffp-contract=off", "-o", so, SRC], check=True)
    L = ctypes.CDLL(so)
    lut = np.ascontiguousarray(load_lut())
    L.k1h_init(lut.ctypes.data_as(ctypes.c_void_p))
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _colours():
    rs = np.random.RandomState(5)
    grey = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 3, 1)
    corners = np.array([[r, g, b] for r in (0, 1, 127, 128, 254, 255) for g in (0, 1, 127, 128, 254, 255)
                        for b in (0, 1, 127, 128, 254, 255)], dtype=np.uint8)
    return np.ascontiguousarray(np.concatenate([grey, corners, rs.randint(0, 256, (400000, 3)).astype(np.uint8)]))


def _oracle_stage_a(x01):
    lab = O.rgb2lab_f32(x01[None], load_lut())[0]
    spc = (lab + np.array([0, 128, 128], dtype=f32)) / np.array([100.0, 255.0, 255.0], dtype=f32)
    return (spc[:, 0] * f32(255)).astype(np.uint8), spc


def test_pass_a_u8_matches_oracle(lib):
    rgb = _colours()
    n = len(rgb)
    l8 = np.empty(n, np.uint8)
    ab = np.empty(n, np.uint32)
    lib.k1h_stage_a_u8(_p(rgb), ctypes.c_long(n), _p(l8), _p(ab))
    ref_l8, spc = _oracle_stage_a(rgb.astype(f32) / f32(255))
    assert np.array_equal(l8, ref_l8)
    # chroma: the kernel keeps the Q14 integers; the oracle's normalised value is (o / 64) / 255
    for ch, o in ((1, ab & 0xffff), (2, ab >> 16)):
        assert np.array_equal((o.astype(f32) * f32(1 / 64.0)) / f32(255), spc[:, ch])


def test_pass_a_f32_matches_oracle(lib):
    rs = np.random.RandomState(6)
    x = rs.rand(300000, 3).astype(f32)
    x[:1000] = rs.randint(0, 3, (1000, 3)).astype(f32) / f32(2)        # gamut corners incl. exactly 1.0
    x[1000:2000] = f32(1) - rs.rand(1000, 3).astype(f32) * f32(1e-5)
    x = np.ascontiguousarray(x)
    n = len(x)
    l8 = np.empty(n, np.uint8)
    ab = np.empty(n, np.uint32)
    lib.k1h_stage_a_f32(_p(x), ctypes.c_long(n), _p(l8), _p(ab))
    ref_l8, spc = _oracle_stage_a(x)
    assert np.array_equal(l8, ref_l8)
    for ch, o in ((1, ab & 0xffff), (2, ab >> 16)):
        assert np.array_equal((o.astype(f32) * f32(1 / 64.0)) / f32(255), spc[:, ch])


@pytest.mark.parametrize("use_table,tail,width", [(1, 0, 8), (0, 0, 8), (0, 1, 1), (2, 0, 8)])
def test_pass_b_matches_oracle(lib, use_table, tail, width):
    """width 8 -> every pixel takes OpenCV's SIMD-body sequence, width 1 -> every pixel is a scalar-tail pixel;
    use_table 2 -> the packed two-pixel (f32x2) sequence the device runs by default."""
    rgb = _colours()
    n = (len(rgb) // 8) * 8
    rgb = rgb[:n]
    l8 = np.empty(n, np.uint8)
    ab = np.empty(n, np.uint32)
    lib.k1h_stage_a_u8(_p(rgb), ctypes.c_long(n), _p(l8), _p(ab))
    dst = np.ascontiguousarray(np.random.RandomState(7).randint(0, 256, n).astype(np.uint8))   # any CLAHE output byte
    out = np.empty((n, 3), f32)
    mean, std = np.array(MEAN, f32), np.array(STD, f32)
    lib.k1h_stage_b(_p(dst), _p(ab), ctypes.c_long(n), use_table, tail, _p(mean), _p(std), _p(out))
    # oracle: normalised Lab with the CLAHE byte in the lightness slot -> denormalise -> LAB2RGB -> normalise
    _, spc = _oracle_stage_a(rgb.astype(f32) / f32(255))
    spc = spc.copy()
    spc[:, 0] = dst.astype(f32) / f32(255)
    lab2 = spc * np.array([100.0, 255.0, 255.0], dtype=f32) - np.array([0, 128, 128], dtype=f32)
    ref = (O.lab2rgb_f32(lab2.reshape(-1, width, 3)).reshape(-1, 3) - mean) / std
    assert np.array_equal(out.view(np.uint32), ref.astype(f32).view(np.uint32))


def test_compressed_lattice_record_equals_trilinear_exhaustively(lib):
    """One 32-byte record per cell (base + first differences + mixed differences of the three channels) must reproduce
    the 8-corner trilinear interpolation for every cell and every 4-bit fraction triple: 35 937 x 4 096 x 3 values."""
    lib.k1h_rec32_check.restype = ctypes.c_long
    bad = lib.k1h_rec32_check(1)
    assert bad == 0, "compressed lattice record: %d mismatches (-1: the table does not fit the record format)" % bad

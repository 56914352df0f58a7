"""Shared helpers of the test-suite: synthetic inputs and golden-fixture access."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LUT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gandtr_b200", "data", "rgb2lab_lut_s16.bin")
MEAN = [0.485, 0.456, 0.406]
STD = [0.229, 0.224, 0.225]


def golden(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def load_lut():
    return np.fromfile(LUT_PATH, dtype="<i2").reshape(33, 33, 33, 3)


def synth_image(seed, h, w, kind):
    rs = np.random.RandomState(seed)
    if kind == "noise":
        return rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([128 + 70 * (np.sin(xx / (37.0 + 5 * c)) + np.cos(yy / (23.0 + 3 * c))) for c in range(3)], -1)
    img = img + rs.normal(0, 8, img.shape)
    if kind == "dark":
        img = 255.0 * (np.clip(img, 0, 255) / 255.0) ** 2.8
    return np.clip(img, 0, 255).astype(np.uint8)


def unit_rows(rs, n, d):
    x = rs.normal(0, 1, (n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)

"""Regenerate gandtr_b200/data/rgb2lab_lut_s16.bin: the 33x33x33x3 int16 lattice table of OpenCV's float
RGB2Lab path (color_lab.cpp, RGB2LabLUT_s16), recovered by probing live cv2 at the lattice colours
(SURVEY.md App. A.1). OpenCV builds the table with softfloat; the float64 analytic restatement in
oracle/clahe_np.py differs in a few dozen entries at .5 rounding boundaries, so the probed table ships.
Run in the build container: python tools/gen_tables.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import clahe_np  # noqa: E402

if __name__ == "__main__":
    import cv2
    lut = clahe_np.probe_rgb2lab_lut_cv2()
    ana = clahe_np.rgb2lab_lut_analytic()
    out = os.path.join(os.path.dirname(__file__), "..", "gandtr_b200", "data", "rgb2lab_lut_s16.bin")
    lut.astype("<i2").tofile(out)
    print("cv2", cv2.__version__, "entries", lut.size, "analytic mismatches", int((lut != ana).sum()), "->", os.path.abspath(out))

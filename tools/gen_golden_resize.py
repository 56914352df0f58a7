"""Golden fixtures of the image-geometry path (K5), produced by Pillow itself through the reference's own call
(`img.thumbnail((imsize, imsize), LANCZOS)` after an optional `img.crop(bbx)`; genericdataset.py:86-97,
datahelpers.py:75-82). Run in the build container:  python tools/gen_golden_resize.py  -> tests/golden/resize.npz"""
import os
import sys

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.util import synth_image  # noqa: E402

CASES = [  # (h, w, kind, imsize, bbx)
    (96, 128, "smooth", 64, None), (61, 83, "noise", 50, None), (150, 45, "dark", 64, None), (90, 120, "smooth", 128, None),
    (129, 259, "smooth", 32, None), (100, 77, "noise", 11, None), (120, 160, "smooth", 128, (30, 40, 130, 110)),
    (64, 64, "noise", 63, None), (32, 500, "smooth", 50, None),
]


def reference_load(img_u8, imsize, bbx):
    img = Image.fromarray(img_u8)
    full = max(img.size)
    if bbx:
        img = img.crop(bbx)
    lanczos = getattr(Image, "LANCZOS", Image.Resampling.LANCZOS)
    img.thumbnail((imsize * max(img.size) / full,) * 2 if bbx else (imsize, imsize), lanczos)
    return np.asarray(img)


if __name__ == "__main__":
    out = {}
    for i, (h, w, kind, imsize, bbx) in enumerate(CASES):
        img = synth_image(900 + i, h, w, kind)
        out["img%d" % i] = img
        out["imsize%d" % i] = np.int64(imsize)
        out["bbx%d" % i] = np.array(bbx if bbx else [], dtype=np.int64)
        out["out%d" % i] = reference_load(img, imsize, bbx)
    import PIL
    out["pillow_version"] = np.array(PIL.__version__)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "resize.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

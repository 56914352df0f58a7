"""Golden fixtures of the image-geometry path (K5), produced by the UNMODIFIED reference: its dataset class
`ImagesFromList` (mdir/external/cirtorch/datasets/genericdataset.py:12-102) loads lossless PNG files of seeded synthetic
images with its own `default_loader`, crops them to the bounding boxes and `imresize`s them (datahelpers.py:75-82) --
i.e. Pillow does the arithmetic, the reference decides what Pillow is asked. When /root/reference is absent the same
call sequence is issued to Pillow directly (`reference_load` below; identical output, asserted when both are available).
Run in the build container:  PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden_resize.py  -> tests/golden/resize.npz"""
import os
import sys

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.util import synth_image  # noqa: E402

CASES = [  # (h, w, kind, imsize, bbx)
    (96, 128, "smooth", 64, None), (61, 83, "noise", 50, None), (150, 45, "dark", 64, None), (90, 120, "smooth", 128, None),
    (129, 259, "smooth", 32, None), (100, 77, "noise", 11, None), (120, 160, "smooth", 128, (30, 40, 130, 110)),
    (120, 160, "smooth", 100, (12.5, 7.4, 140.5, 99.6)),
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


def reference_dataset_load(imgs, imsizes, bbxs):
    """The reference's own ImagesFromList on PNG files (one dataset per imsize, as extract_vectors builds it)."""
    import tempfile
    from oracle import ref_harness
    ref_harness.load_reference()
    sys.path.insert(0, os.path.join(ref_harness.REF_ROOT, "mdir", "external"))
    from cirtorch.datasets.genericdataset import ImagesFromList
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for i, (img, imsize, bbx) in enumerate(zip(imgs, imsizes, bbxs)):
            path = os.path.join(d, "im%02d.png" % i)
            Image.fromarray(img).save(path)
            ds = ImagesFromList(root="", images=[path], imsize=imsize, bbxs=[bbx] if bbx else None, transform=None)
            outs.append(np.asarray(ds[0]))
    return outs


if __name__ == "__main__":
    out = {}
    imgs = [synth_image(900 + i, h, w, kind) for i, (h, w, kind, _, _) in enumerate(CASES)]
    from oracle import ref_harness
    refs = reference_dataset_load(imgs, [c[3] for c in CASES], [c[4] for c in CASES]) if ref_harness.available() else None
    for i, (h, w, kind, imsize, bbx) in enumerate(CASES):
        img = imgs[i]
        out["img%d" % i] = img
        out["imsize%d" % i] = np.int64(imsize)
        out["bbx%d" % i] = np.array(bbx if bbx else [], dtype=np.float64)
        out["out%d" % i] = reference_load(img, imsize, bbx)
        if refs is not None:
            assert refs[i].shape == out["out%d" % i].shape and np.array_equal(refs[i], out["out%d" % i]), i
            out["out%d" % i] = refs[i]
    out["source"] = np.array("reference ImagesFromList" if refs is not None else "Pillow, reference call sequence")
    import PIL
    out["pillow_version"] = np.array(PIL.__version__)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "resize.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

"""TEST INFRASTRUCTURE ONLY (see oracle/__init__ note in DESIGN.md section 2): CPU restatement of the image-size half
of the reference's dataset loader, i.e. what `imresize` (mdir/external/cirtorch/datasets/datahelpers.py:75-82) and
`ImagesFromList.__getitem__` (genericdataset.py:86-97) make Pillow do:

    img.crop(bbx)                                    -> a sub-rectangle
    img.thumbnail((imsize, imsize), Image.LANCZOS)   -> aspect-preserving target size, optional integer box `reduce`
                                                        (reducing_gap = 2.0), then the two-pass 8-bit LANCZOS resample

The arithmetic lives in a third-party dependency that is absent from /root/reference: Pillow (`pillow`, unpinned in
requirements.txt; 12.2.0 in this image). Restated from its published algorithm (src/libImaging/Resample.c: precompute_coeffs,
normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc; src/libImaging/Reduce.c; PIL/Image.py: thumbnail,
resize, _get_safe_box) and PINNED bit-exact against live Pillow in tests/test_oracle_resize.py and against fixtures made
by Pillow itself (tests/golden/resize.npz, tools/gen_golden_resize.py).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos_filter(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def precompute_coeffs(in_size, in0, in1, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the LANCZOS filter.
    in0 / in1 are C floats in Pillow. -> (ksize, bounds[out_size, 2] (xmin, count), kk[out_size, ksize] int32)."""
    in0 = float(np.float32(in0))
    in1 = float(np.float32(in1))
    scale = filterscale = float(np.float32(in1) - np.float32(in0)) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [lanczos_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            kk[xx, x] = int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass(src, bounds, kk, axis):
    """One 8-bit resample pass along `axis` (1 = horizontal, 0 = vertical) of an [h, w, c] uint8 array."""
    src = np.moveaxis(src, axis, 0).astype(np.int64)
    out = np.empty((len(bounds),) + src.shape[1:], dtype=np.uint8)
    for o, (lo, cnt) in enumerate(bounds):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[lo + x] * int(kk[o, x])
        out[o] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resample_lanczos_u8(img, out_w, out_h, box=None):
    """ImagingResample on an [h, w, c] uint8 image: horizontal pass over the rows the vertical pass needs, then the
    vertical pass; a pass is skipped when that axis keeps its size and the box is the whole axis."""
    img = np.asarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    if box is None:
        box = (0, 0, w, h)
    need_h = out_w != w or box[0] != 0 or box[2] != out_w
    need_v = out_h != h or box[1] != 0 or box[3] != out_h
    _, bh, kh = precompute_coeffs(w, box[0], box[2], out_w)
    _, bv, kv = precompute_coeffs(h, box[1], box[3], out_h)
    first = int(bv[0, 0])
    last = int(bv[-1, 0] + bv[-1, 1])
    out = img
    if need_h:
        bv = bv.copy()
        bv[:, 0] -= first
        out = _pass(img[first:last], bh, kh, axis=1)
    if need_v:
        out = _pass(out, bv, kv, axis=0)
    return np.ascontiguousarray(out)


def _division_u32(divider, result_bits):
    max_dividend = (1 << result_bits) * divider
    return int(np.float32(np.float32(1 << 30) * np.float32(4.0)) / np.float32(max_dividend))


def reduce_u8(img, fx, fy):
    """Image.reduce((fx, fy)) on the whole [h, w, c] uint8 image (Reduce.c): box average with a fixed-point reciprocal;
    the ragged last column / row average over the pixels that exist."""
    img = np.asarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    oh, ow = (h + fy - 1) // fy, (w + fx - 1) // fx
    out = np.empty((oh, ow) + img.shape[2:], dtype=np.uint8)
    src = img.astype(np.int64)
    for oy in range(oh):
        y0, y1 = oy * fy, min((oy + 1) * fy, h)
        rows = src[y0:y1].sum(axis=0)
        cs = np.concatenate([np.zeros((1,) + rows.shape[1:], dtype=np.int64), np.cumsum(rows, axis=0)])
        for full, (lo, hi) in ((True, (0, w // fx)), (False, (w // fx, ow))):
            if hi <= lo:
                continue
            xs = np.arange(lo, hi)
            x0, x1 = xs * fx, np.minimum((xs + 1) * fx, w)
            n = (x1 - x0) * (y1 - y0)
            ssum = cs[x1] - cs[x0]
            mult = np.array([_division_u32(int(v), 8) for v in n], dtype=np.int64)
            amend = (n // 2).astype(np.int64)
            shape = (-1,) + (1,) * (ssum.ndim - 1)
            out[oy, lo:hi] = (((ssum + amend.reshape(shape)) * mult.reshape(shape)) >> 24).astype(np.uint8)
    return out


def thumbnail_size(w, h, imsize):
    """Image.thumbnail's target size for the request (imsize, imsize); None = the image is left alone."""
    x = y = math.floor(imsize)
    if x >= w and y >= h:
        return None

    def round_aspect(number, key):
        return max(min(math.floor(number), math.ceil(number), key=key), 1)

    aspect = w / h
    if x / y >= aspect:
        x = round_aspect(y * aspect, key=lambda n: abs(aspect - n / y))
    else:
        y = round_aspect(x / aspect, key=lambda n: 0 if n == 0 else abs(aspect - x / n))
    return x, y


def reduce_factors(w, h, out_w, out_h, reducing_gap=2.0):
    """Image.resize's pre-reduction factors for box = the whole image."""
    return int(w / out_w / reducing_gap) or 1, int(h / out_h / reducing_gap) or 1


def thumbnail_u8(img, imsize):
    """img.thumbnail((imsize, imsize), LANCZOS) on an [h, w, c] uint8 array (a converted, fully loaded image: JPEG draft
    mode does not apply because the reference's pil_loader returns img.convert('RGB'))."""
    img = np.asarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    size = thumbnail_size(w, h, imsize)
    if size is None or size == (w, h):
        return img
    out_w, out_h = size
    fx, fy = reduce_factors(w, h, out_w, out_h)
    box = (0, 0, w, h)
    if fx > 1 or fy > 1:
        img = reduce_u8(img, fx, fy)        # _get_safe_box of the whole image is the whole image
        box = (0.0, 0.0, w / fx, h / fy)
    return resample_lanczos_u8(img, out_w, out_h, box)


def load_resized_u8(img, imsize=None, bbx=None):
    """genericdataset.py:86-97 on a decoded [h, w, 3] uint8 image: optional crop, then thumbnail with the reference's
    bounding-box scale rule."""
    img = np.asarray(img, dtype=np.uint8)
    full = max(img.shape[0], img.shape[1])
    if bbx:
        x0, y0, x1, y1 = [int(round(v)) for v in bbx]     # Image.crop rounds the box; outside pixels are black
        h, w = img.shape[:2]
        out = np.zeros((max(y1 - y0, 0), max(x1 - x0, 0)) + img.shape[2:], dtype=np.uint8)
        sx0, sy0, sx1, sy1 = max(x0, 0), max(y0, 0), min(x1, w), min(y1, h)
        if sx1 > sx0 and sy1 > sy0:
            out[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = img[sy0:sy1, sx0:sx1]
        img = out
    if imsize is not None:
        img = thumbnail_u8(img, imsize * max(img.shape[0], img.shape[1]) / full if bbx else imsize)
    return np.ascontiguousarray(img)

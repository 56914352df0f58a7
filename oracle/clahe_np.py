"""ORACLE (test infrastructure, never imported by gandtr_b200/): NumPy restatement of the
reference transform  `pil2np | apply_clahe:1.0 | totensor | normalize`.

Follows, stage by stage:
  * mdir/components/data/transform/core_transforms.py:76-100   (Pil2Numpy: u8 -> f32 / 255.0)
  * mdir/components/data/transform/functional.py:28-35         (rgb2normspace 'lab')
  * mdir/components/data/transform/functional.py:140-151       (ChannelClahe.apply_clahe)
  * mdir/components/data/transform/functional.py:55-63         (normspace2rgb 'lab')
  * mdir/components/data/transform/functional.py:81-85         (apply_lightness_transform)
  * mdir/components/data/transform/photometric_transforms.py:28-36 (ApplyClahe)
  * mdir/components/data/transform/core_transforms.py:35-70    (ToTensor, Normalize)

The arithmetic inside `cv2.cvtColor` / `cv2.createCLAHE` is third-party (opencv-python,
unpinned in the reference's requirements.txt:3; 4.13.0 in this image) and is restated from its
published algorithm (modules/imgproc/src/color_lab.cpp, clahe.cpp) as characterised in
SURVEY.md App. A.  Parity pin: tests/test_oracle_clahe.py checks this file bit-for-bit against
live cv2 4.13.0 *and* against golden fixtures produced by the unmodified reference
(tools/gen_golden.py).

Everything is IEEE binary32 with one rounding per written operation (NumPy float32 ops do not
contract to FMA).
"""
import numpy as np

f32 = np.float32

LAB_LUT_DIM = 33
LAB_BASE = 16384  # 1 << 14


# --------------------------------------------------------------------------------------------
# Tables
# --------------------------------------------------------------------------------------------

def _srgb_gamma_f64(x):
    x = np.asarray(x, dtype=np.float64)
    return np.where(x <= 0.04045, x / 12.92, np.power((x + 0.055) / 1.055, 2.4))


def rgb2lab_lut_analytic():
    """33x33x33x3 int16 table of color_lab.cpp (initLabTabs, RGB2LabLUT_s16 lattice values),
    computed in float64. cv2 computes it with softfloat (correctly rounded binary32); the two
    agree except where a value falls within float32 rounding of a .5 boundary -- the tests
    compare this against the table probed from live cv2 and report the mismatch count."""
    D65 = np.array([0.950456, 1.0, 1.088754])
    M = np.array([[0.412453, 0.357580, 0.180423],
                  [0.212671, 0.715160, 0.072169],
                  [0.019334, 0.119193, 0.950227]])
    C = M / D65[:, None]
    g = _srgb_gamma_f64(np.arange(LAB_LUT_DIM) / (LAB_LUT_DIM - 1.0))
    R, G, B = np.meshgrid(g, g, g, indexing="ij")
    X = R * C[0, 0] + G * C[0, 1] + B * C[0, 2]
    Y = R * C[1, 0] + G * C[1, 1] + B * C[1, 2]
    Z = R * C[2, 0] + G * C[2, 1] + B * C[2, 2]
    thr = 216.0 / 24389.0
    sc = 841.0 / 108.0
    bias = 16.0 / 116.0

    def fxyz(t):
        return np.where(t > thr, np.cbrt(t), t * sc + bias)

    FX, FY, FZ = fxyz(X), fxyz(Y), fxyz(Z)
    L = np.where(Y > thr, 116.0 * FY - 16.0, 903.3 * Y)
    a = 500.0 * (FX - FY)
    b = 200.0 * (FY - FZ)
    out = np.stack([np.rint(LAB_BASE * L / 100.0),
                    np.rint(LAB_BASE * (a + 128.0) / 256.0),
                    np.rint(LAB_BASE * (b + 128.0) / 256.0)], axis=-1)
    return out.astype(np.int16)


def probe_rgb2lab_lut_cv2():
    """Recover the lattice table from live cv2: at lattice colours all
    trilinear weights but one vanish, so the integer output is the table entry itself."""
    import cv2
    g = (np.arange(LAB_LUT_DIM, dtype=np.float32) / f32(LAB_LUT_DIM - 1))
    R, G, B = np.meshgrid(g, g, g, indexing="ij")
    img = np.stack([R, G, B], axis=-1).reshape(LAB_LUT_DIM * LAB_LUT_DIM, LAB_LUT_DIM, 3).astype(np.float32)
    lab = cv2.cvtColor(img, cv2.COLOR_RGB2LAB).reshape(LAB_LUT_DIM, LAB_LUT_DIM, LAB_LUT_DIM, 3).astype(np.float64)
    out = np.stack([np.rint(lab[..., 0] / 100.0 * LAB_BASE),
                    np.rint((lab[..., 1] + 128.0) / 256.0 * LAB_BASE),
                    np.rint((lab[..., 2] + 128.0) / 256.0 * LAB_BASE)], axis=-1)
    return out.astype(np.int16)


def inv_gamma_spline_tab():
    """1024x4 float32 natural-cubic-spline table of the sRGB inverse gamma
    (color_lab.cpp: sRGBInvGammaTab via splineBuild, GAMMA_TAB_SIZE=1024). SURVEY.md App. A.3."""
    n = 1024
    xi = np.arange(n + 1, dtype=np.float64) / n
    f = np.where(xi <= 0.0031308, xi * 12.92, 1.055 * np.power(xi, 1.0 / 2.4) - 0.055).astype(np.float32)
    tab = np.zeros(n * 4, dtype=np.float32)
    three, four, two, one = f32(3), f32(4), f32(2), f32(1)
    for i in range(1, n):
        t = (f[i + 1] - f[i] * two + f[i - 1]) * three
        l = one / (four - tab[(i - 1) * 4])
        tab[i * 4] = l
        tab[i * 4 + 1] = (t - tab[(i - 1) * 4 + 1]) * l
    cn = f32(0)
    for i in range(n - 1, -1, -1):
        c = tab[i * 4 + 1] - tab[i * 4] * cn
        b = f[i + 1] - f[i] - (cn + c * two) / three
        d = (cn - c) / three
        tab[i * 4] = f[i]
        tab[i * 4 + 1] = b
        tab[i * 4 + 2] = c
        tab[i * 4 + 3] = d
        cn = c
    return tab.reshape(n, 4)


def lab2rgb_coeffs():
    """C[k][j] = float32(XYZ2sRGB_D65[k][j] * whitePt[j]) computed in double (App. A.3)."""
    M = np.array([[3.240479, -1.53715, -0.498535],
                  [-0.969256, 1.875991, 0.041556],
                  [0.055648, -0.204043, 1.057311]], dtype=np.float64)
    wp = np.array([0.950456, 1.0, 1.088754], dtype=np.float64)
    return (M * wp[None, :]).astype(np.float32)


# --------------------------------------------------------------------------------------------
# Stages
# --------------------------------------------------------------------------------------------

def rgb2lab_f32(x, lut):
    """cv2.cvtColor(x, COLOR_RGB2LAB) for float32 HxWx3 in [0,1] (integer LUT + trilinear path)."""
    x = np.clip(np.asarray(x, dtype=np.float32), f32(0), f32(1))
    c = np.rint(x * f32(LAB_BASE)).astype(np.int32)          # cvRound = round-half-even
    t = c >> 9
    fr = (c >> 5) & 15
    lut = lut.astype(np.int64)
    acc = np.zeros(x.shape[:2] + (3,), dtype=np.int64)
    for dx in (0, 1):
        wx = fr[..., 0] if dx else 16 - fr[..., 0]
        ix = np.minimum(t[..., 0] + dx, LAB_LUT_DIM - 1)
        for dy in (0, 1):
            wy = fr[..., 1] if dy else 16 - fr[..., 1]
            iy = np.minimum(t[..., 1] + dy, LAB_LUT_DIM - 1)
            for dz in (0, 1):
                wz = fr[..., 2] if dz else 16 - fr[..., 2]
                iz = np.minimum(t[..., 2] + dz, LAB_LUT_DIM - 1)
                w = (wx * wy * wz).astype(np.int64)
                acc += lut[ix, iy, iz] * w[..., None]
    out = (acc + 2048) >> 12
    of = out.astype(np.float32) * (f32(1.0) / f32(LAB_BASE))
    L = of[..., 0] * f32(100.0)
    a = of[..., 1] * f32(256.0) - f32(128.0)
    b = of[..., 2] * f32(256.0) - f32(128.0)
    return np.stack([L, a, b], axis=-1)


def _reflect101(i, n):
    i = np.asarray(i)
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def clahe_u8(src, clip_limit=1.0, grid=8):
    """cv2.createCLAHE(clipLimit, (grid, grid)).apply(src) for uint8 HxW (clahe.cpp). App. A.2."""
    src = np.asarray(src, dtype=np.uint8)
    H, W = src.shape
    if W % grid == 0 and H % grid == 0:
        ext = src
    else:
        eh = H + (grid - H % grid)
        ew = W + (grid - W % grid)
        yy = _reflect101(np.arange(eh), H)
        xx = _reflect101(np.arange(ew), W)
        ext = src[yy][:, xx]
    th, tw = ext.shape[0] // grid, ext.shape[1] // grid
    area = th * tw
    lut_scale = f32(255.0) / f32(area)
    clip = 0
    if clip_limit > 0.0:
        clip = max(int(clip_limit * area / 256), 1)
    luts = np.zeros((grid, grid, 256), dtype=np.uint8)
    for ty in range(grid):
        for tx in range(grid):
            tile = ext[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            hist = np.bincount(tile.ravel(), minlength=256).astype(np.int64)
            if clip > 0:
                clipped = int(np.maximum(hist - clip, 0).sum())
                hist = np.minimum(hist, clip)
                batch = clipped // 256
                resid = clipped - batch * 256
                hist += batch
                if resid:
                    step = max(256 // resid, 1)
                    i = 0
                    while i < 256 and resid > 0:
                        hist[i] += 1
                        i += step
                        resid -= 1
            cs = np.cumsum(hist).astype(np.float32) * lut_scale
            luts[ty, tx] = np.clip(np.rint(cs), 0, 255).astype(np.uint8)
    inv_tw = f32(1.0) / f32(tw)
    inv_th = f32(1.0) / f32(th)
    xs = np.arange(W, dtype=np.float32) * inv_tw - f32(0.5)
    tx1 = np.floor(xs).astype(np.int32)
    xa = (xs - tx1.astype(np.float32)).astype(np.float32)
    xa1 = f32(1.0) - xa
    tx2 = np.minimum(tx1 + 1, grid - 1)
    tx1 = np.maximum(tx1, 0)
    ys = np.arange(H, dtype=np.float32) * inv_th - f32(0.5)
    ty1 = np.floor(ys).astype(np.int32)
    ya = (ys - ty1.astype(np.float32)).astype(np.float32)
    ya1 = f32(1.0) - ya
    ty2 = np.minimum(ty1 + 1, grid - 1)
    ty1 = np.maximum(ty1, 0)
    v = src.astype(np.int64)
    l11 = luts[ty1[:, None], tx1[None, :], v].astype(np.float32)
    l12 = luts[ty1[:, None], tx2[None, :], v].astype(np.float32)
    l21 = luts[ty2[:, None], tx1[None, :], v].astype(np.float32)
    l22 = luts[ty2[:, None], tx2[None, :], v].astype(np.float32)
    res = (l11 * xa1[None, :] + l12 * xa[None, :]) * ya1[:, None] + \
          (l21 * xa1[None, :] + l22 * xa[None, :]) * ya[:, None]
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


def lab2rgb_f32(lab, tab=None, C=None):
    """cv2.cvtColor(lab, COLOR_LAB2RGB) on float32 HxWx3 (cv2 4.13.0 pip wheel). App. A.3:
    the first (W//8)*8 pixels of every row take the SIMD body (reciprocal multiplies), the
    remaining W%8 pixels the scalar tail (true divisions, different association)."""
    lab = np.asarray(lab, dtype=np.float32)
    tab = inv_gamma_spline_tab() if tab is None else tab
    C = lab2rgb_coeffs() if C is None else C
    H, W, _ = lab.shape
    wb = (W // 8) * 8
    out = np.empty_like(lab)
    c16 = f32(16.0) / f32(116.0)
    fth = f32(6.0) / f32(29.0)

    def spline(lin):
        x = np.clip(lin, f32(0), f32(1)) * f32(1024.0)
        ix = np.clip(x.astype(np.int32), 0, 1023)
        x = x - ix.astype(np.float32)
        t = tab[ix]
        return ((t[..., 3] * x + t[..., 2]) * x + t[..., 1]) * x + t[..., 0]

    if wb:
        L, a, b = lab[:, :wb, 0], lab[:, :wb, 1], lab[:, :wb, 2]
        r903, r116, r500, r200, r7787 = (f32(1.0) / f32(903.3), f32(1.0) / f32(116.0), f32(1.0) / f32(500.0),
                                         f32(1.0) / f32(200.0), f32(1.0) / f32(7.787))
        ylo = L * r903
        fylo = ylo * f32(7.787) + c16
        fyhi = (L + f32(16.0)) * r116
        yhi = (fyhi * fyhi) * fyhi
        lo = L <= f32(8.0)
        y = np.where(lo, ylo, yhi)
        fy = np.where(lo, fylo, fyhi)
        fx = a * r500 + fy
        fz = fy - b * r200

        def g(f):
            return np.where(f <= fth, (f - c16) * r7787, (f * f) * f)

        X, Z = g(fx), g(fz)
        for k in range(3):
            lin = C[k, 0] * X + (C[k, 1] * y + C[k, 2] * Z)
            out[:, :wb, k] = spline(lin)
    if wb < W:
        L, a, b = lab[:, wb:, 0], lab[:, wb:, 1], lab[:, wb:, 2]
        ylo = L / f32(903.3)
        fylo = ylo * f32(7.787) + c16
        fyhi = (L + f32(16.0)) / f32(116.0)
        yhi = (fyhi * fyhi) * fyhi
        lo = L <= f32(8.0)
        y = np.where(lo, ylo, yhi)
        fy = np.where(lo, fylo, fyhi)
        fx = a / f32(500.0) + fy
        fz = fy - b / f32(200.0)

        def g2(f):
            return np.where(f <= fth, (f - c16) / f32(7.787), (f * f) * f)

        X, Z = g2(fx), g2(fz)
        for k in range(3):
            lin = (C[k, 0] * X + C[k, 1] * y) + C[k, 2] * Z
            out[:, wb:, k] = spline(lin)
    return out


def apply_clahe_rgb_f32(img, lut, clip_limit=1.0, grid=8, tab=None, C=None):
    """ApplyClahe(clip, grid, 'lab') on a float32 HxWx3 image in [0,1]
    (photometric_transforms.py:28-36 -> functional.py:154-161,81-85)."""
    img = np.asarray(img, dtype=np.float32)
    lab = rgb2lab_f32(img, lut)
    spc = (lab + np.array([0, 128, 128], dtype=np.float32)) / np.array([100.0, 255.0, 255.0], dtype=np.float32)
    L8 = (spc[..., 0] * f32(255)).astype(np.uint8)                       # truncation (functional.py:148)
    L8c = clahe_u8(L8, clip_limit, grid)
    spc[..., 0] = L8c.astype(np.float32) / f32(255.0)
    lab2 = spc * np.array([100.0, 255.0, 255.0], dtype=np.float32) - np.array([0, 128, 128], dtype=np.float32)
    return lab2rgb_f32(lab2, tab, C)


def transform_u8(img_u8, lut, mean, std, clip_limit=1.0, grid=8, tab=None, C=None):
    """Full `pil2np | apply_clahe:clip | totensor | normalize` on a uint8 HxWx3 image.
    Returns float32 3xHxW."""
    x = np.asarray(img_u8, dtype=np.uint8).astype(np.float32) / f32(255.0)   # core_transforms.py:83
    rgb = apply_clahe_rgb_f32(x, lut, clip_limit, grid, tab, C)
    chw = np.ascontiguousarray(rgb.transpose(2, 0, 1))
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    return (chw - m) / s                                                  # sub then true division


def clahe_post_f32(x_chw_norm, lut, meanstd, clip_limit=1.0, grid=8, tab=None, C=None):
    """ClahePost.postprocess on one normalised float32 CHW image (wrapper.py:334-348)."""
    m = np.asarray(meanstd[0], dtype=np.float32)[:, None, None]
    s = np.asarray(meanstd[1], dtype=np.float32)[:, None, None]
    t = np.asarray(x_chw_norm, dtype=np.float32) * s + m
    img = t.transpose(1, 2, 0)
    rgb = apply_clahe_rgb_f32(img, lut, clip_limit, grid, tab, C)
    return (np.ascontiguousarray(rgb.transpose(2, 0, 1)) - m) / s

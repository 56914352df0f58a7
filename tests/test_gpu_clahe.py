"""K1 parity (through the C ABI): bit-exact against the oracle and the reference's golden outputs."""
import ctypes
import numpy as np
import pytest
import torch

from oracle import clahe_np as O
from tests.util import MEAN, STD, golden, load_lut, synth_image

pytestmark = pytest.mark.gpu


def _run_u8(imgs, clip=1.0, grid=8):
    from gandtr_b200 import _lib
    x = torch.from_numpy(np.stack(imgs)).cuda()
    out = _lib.clahe_u8(x, MEAN, STD, clip, grid)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _assert_bits(a, b, what):
    assert a.shape == b.shape and a.dtype == b.dtype == np.float32
    bad = int((a.view(np.uint32) != b.view(np.uint32)).sum())
    assert bad == 0, "%s: %d of %d floats differ (max abs %g)" % (what, bad, a.size, np.abs(a - b).max())


def test_reference_golden_bit_exact():
    g = golden("clahe_transform.npz")
    n = len([k for k in g.files if k.startswith("img")])
    for i in range(n):
        out = _run_u8([g["img%d" % i]])[0]
        _assert_bits(out, g["out%d" % i], "golden case %d %s" % (i, g["img%d" % i].shape))


@pytest.mark.parametrize("kind,h,w", [("smooth", 96, 128), ("noise", 64, 64), ("dark", 120, 160), ("smooth", 61, 77),
                                      ("noise", 37, 53), ("smooth", 100, 130), ("smooth", 8, 8), ("smooth", 72, 100),
                                      ("dark", 50, 1100), ("smooth", 683, 1024), ("noise", 9, 64), ("smooth", 20, 12),
                                      ("noise", 75, 36), ("smooth", 41, 2044)])
def test_oracle_bit_exact_shapes(kind, h, w):
    lut = load_lut()
    img = synth_image(100 + h + w, h, w, kind)
    out = _run_u8([img])[0]
    _assert_bits(out, O.transform_u8(img, lut, MEAN, STD), "%s %dx%d" % (kind, h, w))


def test_batch_and_clip_limits():
    lut = load_lut()
    imgs = [synth_image(200 + i, 80, 120, k) for i, k in enumerate(["smooth", "noise", "dark"])]
    for clip in (1.0, 4.0, 0.0):
        out = _run_u8(imgs, clip=clip)
        for i, img in enumerate(imgs):
            _assert_bits(out[i], O.transform_u8(img, lut, MEAN, STD, clip_limit=clip), "clip %g img %d" % (clip, i))


@pytest.mark.parametrize("h,w", [(37, 53), (45, 81), (64, 130), (33, 1027)])
def test_batches_of_odd_widths_bit_exact(h, w):
    """Widths that are not a multiple of 4: image k of a batch starts at an arbitrary byte (uint8 input) and at a float
    that is not 16-byte aligned (planar output); rows are realigned by funnel shifts, stores fall back per row/channel."""
    lut = load_lut()
    imgs = [synth_image(700 + i + w, h, w, "noise" if i == 1 else "smooth") for i in range(5)]
    out = _run_u8(imgs)
    for i, img in enumerate(imgs):
        _assert_bits(out[i], O.transform_u8(img, lut, MEAN, STD), "odd batch %dx%d img %d" % (h, w, i))


def test_clahe_post_float_variant_odd_width_bit_exact():
    from gandtr_b200 import _lib
    lut = load_lut()
    rs = np.random.RandomState(12)
    x = rs.rand(2, 3, 41, 67).astype(np.float32) * 2 - 1          # normalised with mean = std = 0.5
    y = _lib.clahe_f32(torch.from_numpy(x).cuda(), [0.5] * 3, [0.5] * 3, MEAN, STD, 1.0, 8).cpu().numpy()
    for i in range(2):
        t = (x[i] * np.float32(0.5) + np.float32(0.5)).transpose(1, 2, 0)
        ref = O.apply_clahe_rgb_f32(t, lut)
        m, s = np.asarray(MEAN, np.float32)[:, None, None], np.asarray(STD, np.float32)[:, None, None]
        _assert_bits(y[i], (np.ascontiguousarray(ref.transpose(2, 0, 1)) - m) / s, "clahepost odd width %d" % i)


def test_full_size_image_bit_exact():
    """BASELINE config size (1024x768)."""
    lut = load_lut()
    img = synth_image(7, 768, 1024, "smooth")
    out = _run_u8([img])[0]
    _assert_bits(out, O.transform_u8(img, lut, MEAN, STD), "768x1024")


def test_clahe_post_float_variant_bit_exact():
    from gandtr_b200 import _lib
    g = golden("clahe_post.npz")
    x0 = torch.from_numpy(g["x0"][None]).cuda()
    y0 = _lib.clahe_f32(x0, [0.5] * 3, [0.5] * 3, [0.5] * 3, [0.5] * 3, 1.0, 8)[0].cpu().numpy()
    _assert_bits(y0, g["y0"], "clahepost 0.5/0.5")
    x1 = torch.from_numpy(g["x1"][None]).cuda()
    y1 = _lib.clahe_f32(x1, MEAN, STD, MEAN, STD, 4.0, 8)[0].cpu().numpy()
    _assert_bits(y1, g["y1"], "clahepost imagenet clip 4")


def test_idempotent_properties_full_batch():
    """Size-independent properties at bench scale: determinism and batch independence."""
    imgs = [synth_image(300 + i, 768, 1024, "smooth" if i % 4 else "noise") for i in range(4)]
    a = _run_u8(imgs)
    b = _run_u8(imgs[::-1])[::-1]
    assert np.array_equal(a, b)
    assert np.isfinite(a).all()


def test_errors_are_loud():
    from gandtr_b200 import _lib
    with pytest.raises(_lib.GdtError):
        _lib.clahe_u8(torch.zeros((1, 8, 8, 3), dtype=torch.uint8), MEAN, STD)      # CPU tensor: no fallback
    with pytest.raises(_lib.GdtError):
        _lib.clahe_u8(torch.zeros((1, 8, 8, 3), dtype=torch.float32).cuda(), MEAN, STD)


@pytest.mark.parametrize("std", [0.229, 0.224, 0.225, 0.5, 1.0, 0.3333333])
def test_divider_free_normalisation_is_ieee_division_exhaustive(std):
    """K1's `(x - mean) / std` uses a divider-free sequence: check EVERY float a with 2^-30 <= |a| <= 8 against a / std."""
    import ctypes
    from gandtr_b200 import _lib
    lib = _lib.load()
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    lo = np.float32(2.0 ** -30).view(np.uint32)
    hi = np.float32(8.0).view(np.uint32)
    _lib.check(lib.gdt_debug_div_check(ctypes.c_float(std), int(lo), int(hi), ctypes.c_void_p(cnt.data_ptr()),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "gdt_debug_div_check")
    assert int(cnt.item()) == 0


@pytest.mark.parametrize("grid,h,w", [(4, 64, 96), (16, 128, 160), (16, 100, 130), (2, 33, 47), (1, 40, 40)])
def test_other_tile_grids_bit_exact(grid, h, w):
    """grid <= 8 uses the 8-byte LUT rows, 9..16 the 16-byte rows; non-divisible sizes take the reflect-padded path."""
    lut = load_lut()
    img = synth_image(400 + grid + h, h, w, "smooth")
    out = _run_u8([img], clip=2.0, grid=grid)[0]
    _assert_bits(out, O.transform_u8(img, lut, MEAN, STD, clip_limit=2.0, grid=grid), "grid %d %dx%d" % (grid, h, w))


def test_normalisation_falls_back_to_hardware_division_for_awkward_std():
    """A std whose significand is all ones is outside the proven range of the divider-free sequence: K1 must switch to
    IEEE division and stay bit-exact."""
    from gandtr_b200 import _lib
    lut = load_lut()
    img = synth_image(5, 48, 64, "smooth")
    std = [float(np.float32(0.99999994)), 0.224, float(np.uint32(0x3E7FFFFF).view(np.float32))]
    out = _lib.clahe_u8(torch.from_numpy(img[None]).cuda(), MEAN, std)[0].cpu().numpy()
    _assert_bits(out, O.transform_u8(img, lut, MEAN, std), "awkward std")


def test_every_pipe_variant_is_bit_identical():
    """K1's A/B switches (texture pipe vs LSU / shared memory, table vs recomputed lightness half) never change a bit."""
    from gandtr_b200 import _lib
    lib = _lib.load()
    lut = load_lut()
    img = synth_image(11, 96, 128, "smooth")
    ref = O.transform_u8(img, lut, MEAN, STD)
    try:
        for rec32, persist, pack, div1, cf in ((1, 1, 0, 1, 1), (1, 1, 0, 1, 0), (1, 1, 0, 0, 1), (0, 0, 0, 1, 1), (1, 0, 1, 1, 1),
                                               (1, 1, 1, 1, 0), (0, 1, 1, 0, 1)):
            _lib.check(lib.gdt_debug_k1_chroma_f(cf), "chroma_f")
            _lib.check(lib.gdt_debug_k1_div1(div1), "div1")
            _lib.check(lib.gdt_debug_k1_rec32(rec32), "rec32")
            _lib.check(lib.gdt_debug_k1_persist(2 * persist), "persist")
            _lib.check(lib.gdt_debug_k1_pack(pack), "pack")
            for chroma_a, texab, occ_a in ((-1, 0, 4), (1, 1, 4), (1, 0, 4), (1, 1, 6), (0, 0, 4), (0, 0, 6), (0, 2, 4), (0, 4, 4)):
                for spltex in (0, 1):
                    for fytex in (0, 1):
                        _lib.check(lib.gdt_debug_k1_config(texab, spltex, fytex, chroma_a, occ_a), "gdt_debug_k1_config")
                        what = "variant %d %d %d %d %d rec32=%d persist=%d pack=%d div1=%d chroma_f=%d" % (texab, spltex, fytex, chroma_a, occ_a, rec32, persist, pack, div1, cf)
                        _assert_bits(_run_u8([img])[0], ref, what)
                        _assert_bits(_run_u8([img[:61, :77]])[0], O.transform_u8(img[:61, :77], lut, MEAN, STD), "generic path, " + what)
    finally:
        _lib.k1_config_default()


def test_one_step_division_is_verified_for_the_reference_constants():
    """gdt_init tries every numerator the normalisation can see against IEEE division with ONE correction step; the
    ImageNet std values and 0.5 must pass on this device (otherwise the persistent pass B silently keeps two steps)."""
    from gandtr_b200 import _lib
    lib = _lib.load()
    _lib.clahe_u8(torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device="cuda"), MEAN, STD)     # forces gdt_init
    got = {s: lib.gdt_debug_k1_div1_verified(ctypes.c_float(s)) for s in (0.229, 0.224, 0.225, 0.5)}
    assert all(v in (0, 1) for v in got.values()), got
    print("one-step division verified:", got)
    assert got[0.5] == 1

"""K5 parity (through the C ABI): crop + LANCZOS thumbnail on the device, bit-exact against the oracle, against fixtures
made by Pillow, and against live Pillow when it is installed."""
import numpy as np
import pytest
import torch

from oracle import resize_np as R
from tests.util import golden, synth_image

pytestmark = pytest.mark.gpu


def _dev(img, imsize, bbx=None):
    from gandtr_b200.loader import DeviceImageLoader
    out = DeviceImageLoader(imsize=imsize, device="cuda").resize(torch.from_numpy(img), bbx=bbx)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_pillow_golden_bit_exact():
    g = golden("resize.npz")
    n = len([k for k in g.files if k.startswith("img")])
    for i in range(n):
        bbx = tuple(float(v) for v in g["bbx%d" % i]) or None
        out = _dev(g["img%d" % i], int(g["imsize%d" % i]), bbx)
        ref = g["out%d" % i]
        assert out.shape == ref.shape and np.array_equal(out, ref), "golden case %d" % i


@pytest.mark.parametrize("h,w,imsize", [(60, 80, 40), (61, 83, 50), (100, 37, 64), (128, 128, 64), (33, 31, 32), (480, 640, 512),
                                        (50, 70, 80), (400, 90, 33), (257, 515, 64), (777, 333, 50), (90, 1100, 60), (3, 700, 64),
                                        (700, 2, 64), (64, 64, 63), (301, 299, 150.5)])
def test_oracle_bit_exact_shapes(h, w, imsize):
    img = synth_image(500 + h + w, h, w, "noise" if (h + w) % 3 == 0 else "smooth")
    out = _dev(img, imsize)
    ref = R.thumbnail_u8(img, imsize)
    assert out.shape == ref.shape and np.array_equal(out, ref)


def test_extremes_and_crop():
    a = np.full((90, 120, 3), 255, np.uint8)
    a[::7, ::5] = 0                                            # LANCZOS overshoot on both sides of the uint8 range
    assert np.array_equal(_dev(a, 64), R.thumbnail_u8(a, 64))
    assert np.array_equal(_dev(255 - a, 64), R.thumbnail_u8(255 - a, 64))
    img = synth_image(77, 300, 400, "smooth")
    for bbx in [(30, 40, 330, 250), (1, 1, 399, 299), (100, 0, 103, 300), (136.5, 34.1, 348.5, 255.7), (-10, -5, 200, 310)]:
        assert np.array_equal(_dev(img, 256, bbx), R.load_resized_u8(img, 256, bbx)), bbx
    assert np.array_equal(_dev(img, None, (10, 20, 200, 220)), img[20:220, 10:200])      # crop only


def test_full_size_photo_matches_live_pillow():
    """Config-sized case: a 3000 x 2000 'photo' to imsize 1024 (scale 2.93, no pre-reduction) and a 4100-wide one
    (pre-reduction by 2), compared with Pillow itself."""
    PIL = pytest.importorskip("PIL")
    from PIL import Image
    for h, w in [(2000, 3000), (2733, 4100)]:
        img = synth_image(9, h, w, "smooth")
        im = Image.fromarray(img)
        im.thumbnail((1024, 1024), getattr(Image, "LANCZOS", Image.Resampling.LANCZOS))
        out = _dev(img, 1024)
        assert out.shape == np.asarray(im).shape and np.array_equal(out, np.asarray(im))


def test_loader_matches_host_load_image(tmp_path):
    """DeviceImageLoader.load == extract.load_image (the reference's host path) on files, PIL images and arrays."""
    from PIL import Image
    from gandtr_b200.extract import load_image
    from gandtr_b200.loader import DeviceImageLoader
    img = synth_image(31, 240, 320, "smooth")
    path = str(tmp_path / "a.png")
    Image.fromarray(img).save(path)
    ld = DeviceImageLoader(imsize=128, device="cuda")
    for item, bbx in [(path, None), (path, (10, 20, 300, 200)), (Image.fromarray(img), None)]:
        ref = load_image(item if not isinstance(item, Image.Image) else item.copy(), 128, bbx)
        assert np.array_equal(ld.load(item, bbx=bbx).cpu().numpy(), ref)


def test_nvjpeg_decode_path_is_close_to_libjpeg(tmp_path):
    """decode='nvjpeg' hands the JPEG bit streams to the nvJPEG library (gdt_jpeg_decode_batch). It is NOT bit-identical to the
    reference's libjpeg decode -- the tolerance here is the documented difference: mean |diff| < 1 grey level after the
    thumbnail, same shape."""
    from PIL import Image
    from gandtr_b200.extract import load_image
    from gandtr_b200.loader import DeviceImageLoader
    img = synth_image(41, 480, 640, "smooth")
    path = str(tmp_path / "a.jpg")
    Image.fromarray(img).save(path, quality=92)
    from gandtr_b200 import _lib
    if not _lib.jpeg_available():                      # no CUDA toolkit libraries on this machine
        pytest.skip("libnvjpeg could not be loaded")
    ld = DeviceImageLoader(imsize=256, device="cuda", decode="nvjpeg")
    out = ld.load(path).cpu().numpy()
    print("nvJPEG backend of the last batch (1 = hardware engines, 2 = default):", _lib.load().gdt_debug_jpeg_last_backend())
    ref = load_image(path, 256, None)
    assert out.shape == ref.shape
    assert np.abs(out.astype(np.int32) - ref.astype(np.int32)).mean() < 1.0
    # batched form: one GPU decode call for the JPEG files of the list, one K5 launch pair per decoded size; a PNG and an
    # array in the same list take the per-item route
    path2, png = str(tmp_path / "b.jpg"), str(tmp_path / "c.png")
    Image.fromarray(img[::-1].copy()).save(path2, quality=92)
    Image.fromarray(img[:200, :300].copy()).save(png)
    outs = ld.load_batch([path, png, path2])
    assert np.array_equal(outs[0].cpu().numpy(), out)
    assert np.array_equal(outs[1].cpu().numpy(), load_image(png, 256, None))          # PIL decode + K5: bit-exact
    assert np.array_equal(outs[2].cpu().numpy(), ld.load(path2).cpu().numpy())


def test_errors_are_loud():
    from gandtr_b200 import _lib
    from gandtr_b200.loader import DeviceImageLoader
    with pytest.raises(_lib.GdtError):
        DeviceImageLoader(imsize=64, device="cpu")
    ld = DeviceImageLoader(imsize=64, device="cuda")
    with pytest.raises(_lib.GdtError):
        ld.resize(torch.zeros((8, 8, 3), dtype=torch.float32))
    plan = _lib.ResizePlan(80, 60, 40, "cuda")
    with pytest.raises(_lib.GdtError):
        _lib.resize_u8(plan, torch.zeros((61, 80, 3), dtype=torch.uint8, device="cuda"))


@pytest.mark.parametrize("h,w,imsize,n", [(300, 400, 128, 5), (257, 515, 64, 3), (480, 640, 512, 2), (90, 1100, 60, 4),
                                          (1200, 1600, 200, 3), (61, 83, 50, 37)])
def test_batched_launch_equals_per_image_and_bytewise_kernels(h, w, imsize, n):
    """gdt_resize_u8_batch: n images of one geometry in one launch per pass (incl. more than one 32-image chunk) give the
    per-image results, the dp4a kernels (coefficients split into byte planes) give the byte-wise kernels' bits, and both
    equal the oracle. Sources with different row strides (crop views) may share a batch."""
    from gandtr_b200 import _lib
    from gandtr_b200.loader import DeviceImageLoader
    lib = _lib.load()
    imgs = [synth_image(900 + i, h, w, "noise" if i % 2 else "smooth") for i in range(n)]
    ld = DeviceImageLoader(imsize=imsize, device="cuda")
    dev = [torch.from_numpy(im).cuda() for im in imgs]
    wide = torch.zeros((h, w + 7, 3), dtype=torch.uint8, device="cuda")            # image 0 as a view with a longer row stride
    wide[:, 3:3 + w] = dev[0]
    dev[0] = wide[:, 3:3 + w]
    batch = ld.resize_batch(dev)
    assert tuple(batch.shape[:1]) == (n,) and batch.is_contiguous()
    for i in range(n):
        ref = R.thumbnail_u8(imgs[i], imsize)
        assert np.array_equal(batch[i].cpu().numpy(), ref), "image %d" % i
    try:
        _lib.check(lib.gdt_debug_k5_bytewise(1), "bytewise")
        slow = ld.resize_batch(dev)
    finally:
        _lib.check(lib.gdt_debug_k5_bytewise(0), "bytewise")
    assert torch.equal(slow, batch)
    assert torch.equal(ld.resize(dev[1]), batch[1])
    # sources that are all 4-byte aligned (base and row stride) take the planar horizontal kernel (resize_h5_kernel);
    # with it switched off the interleaved dp4a kernel gives the same bits
    if n > 1:
        fast = ld.resize_batch(dev[1:])
        assert torch.equal(fast, batch[1:])
        try:
            _lib.check(lib.gdt_debug_k5_planar(0), "planar")
            assert torch.equal(ld.resize_batch(dev[1:]), fast)
        finally:
            _lib.check(lib.gdt_debug_k5_planar(1), "planar")

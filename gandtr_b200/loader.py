"""Device-side image geometry of the dataset loader -- SURVEY 8(f) N1. Replaces, on the GPU, what the reference's
DataLoader workers do per image on the host (`ImagesFromList.__getitem__`, mdir/external/cirtorch/datasets/
genericdataset.py:66-102; `imresize`, datahelpers.py:75-82): bounding-box crop and the LANCZOS `thumbnail` to `imsize`,
bit-identical to Pillow (K5, csrc/resize_sm100.cu). Decoding stays with the reference's own decoder (PIL) by default, so
the pixels entering K5 -- and therefore K1's output -- are exactly the reference's; `decode="nvjpeg"` hands the JPEG bit
streams to the GPU instead (gdt_jpeg_decode_batch: the nvJPEG library on the hardware JPEG engines, batched; library
plumbing, NOT bit-identical to libjpeg).

    loader = DeviceImageLoader(imsize=1024, device="cuda")
    img = loader.load(path_or_pil_or_array, bbx=None)      # uint8 CUDA tensor [h, w, 3], what load_image returns on the host
"""
import os

import numpy as np
import torch

from . import _lib
from .extract import crop_like_pil

__all__ = ["DeviceImageLoader", "thumbnail_size"]


def thumbnail_size(w, h, imsize):
    """(out_w, out_h) of `Image.thumbnail((imsize, imsize))` on a w x h image."""
    ow, oh, _, _, _ = _lib.thumbnail_geometry(w, h, imsize)
    return ow, oh


class DeviceImageLoader:
    def __init__(self, imsize=None, device=None, decode="pil", max_plans=256):
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise _lib.GdtError("DeviceImageLoader needs a CUDA device (gandtr_b200 has no CPU path)")
        if decode not in ("pil", "nvjpeg"):
            raise ValueError("decode must be 'pil' or 'nvjpeg'")
        if decode == "nvjpeg" and not _lib.jpeg_available():
            raise _lib.GdtError("decode='nvjpeg' needs libnvjpeg.so (CUDA toolkit); it could not be loaded")
        self.imsize = imsize
        self.decode = decode
        self.max_plans = max_plans
        self._plans = {}

    # -- decoding -----------------------------------------------------------------------------------
    def _decode(self, item):
        """-> uint8 [h, w, 3] tensor, on the device for nvjpeg, pinned host memory otherwise."""
        from PIL import Image
        if isinstance(item, torch.Tensor):
            return item
        if isinstance(item, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(item))
        if isinstance(item, Image.Image):
            return torch.from_numpy(np.asarray(item.convert("RGB")).copy())
        if self.decode == "nvjpeg" and str(item).lower().endswith((".jpg", ".jpeg")):
            with open(item, "rb") as f:
                return _lib.jpeg_decode_batch([f.read()], self.device)[0]
        with open(item, "rb") as f:                       # pil_loader (datahelpers.py:20-27): open + convert('RGB')
            return torch.from_numpy(np.asarray(Image.open(f).convert("RGB")).copy())

    def _plan(self, w, h, imsize):
        key = (w, h, float(imsize))
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= self.max_plans:
                self._plans.pop(next(iter(self._plans)))
            plan = self._plans[key] = _lib.ResizePlan(w, h, imsize, self.device)
        return plan

    # -- geometry -----------------------------------------------------------------------------------
    def resize(self, img, imsize=None, bbx=None):
        """img: decoded uint8 [h, w, 3] tensor (host or device) -> uint8 CUDA [h', w', 3]: crop to `bbx`
        (x0, y0, x1, y1), then thumbnail; the bounding-box scale rule is genericdataset.py:93-97."""
        imsize = self.imsize if imsize is None else imsize
        if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
            raise _lib.GdtError("DeviceImageLoader: expected a uint8 [h, w, 3] image, got %s %s" % (img.dtype, tuple(img.shape)))
        if not img.is_cuda:
            img = (img if img.is_pinned() else img.pin_memory()).to(self.device, non_blocking=True)
        full = max(img.shape[0], img.shape[1])
        if bbx:
            img = crop_like_pil(img, bbx)                 # inside the image: a view, K5 reads it through the parent's row stride
        if imsize is None:
            return img.contiguous()
        h, w = int(img.shape[0]), int(img.shape[1])
        size = imsize * max(w, h) / full if bbx else imsize
        return _lib.resize_u8(self._plan(w, h, size), img)

    def resize_batch(self, imgs, imsize=None):
        """Decoded uint8 [h, w, 3] tensors of ONE size (host or device, no bounding boxes) -> uint8 CUDA [n, h', w', 3]:
        the whole list goes through K5 in one launch per pass (gdt_resize_u8_batch)."""
        imsize = self.imsize if imsize is None else imsize
        dev_imgs = []
        for img in imgs:
            if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3 or img.shape != imgs[0].shape:
                raise _lib.GdtError("DeviceImageLoader.resize_batch: uint8 [h, w, 3] images of one size expected")
            dev_imgs.append(img if img.is_cuda else (img if img.is_pinned() else img.pin_memory()).to(self.device, non_blocking=True))
        if imsize is None:
            return torch.stack([i.contiguous() for i in dev_imgs])
        h, w = int(dev_imgs[0].shape[0]), int(dev_imgs[0].shape[1])
        return _lib.resize_u8_batch(self._plan(w, h, imsize), dev_imgs)

    def crop_only(self, img, bbx=None):
        """Arrays are cropped but never resized by the reference (datahelpers.py:76-79)."""
        if not img.is_cuda:
            img = (img if img.is_pinned() else img.pin_memory()).to(self.device, non_blocking=True)
        if bbx:
            img = crop_like_pil(img, bbx)
        return img.contiguous()

    def load_batch(self, items, imsize=None):
        """Several files at once -> list of uint8 CUDA [h', w', 3] thumbnails. With decode="nvjpeg" the JPEG files of the
        list are handed to the GPU decoder in ONE batched call (gdt_jpeg_decode_batch; pass >= 50 files per call: from
        there on nvJPEG decodes the Huffman streams on the GPU too -- 146 against 44 photos/s of 7 MP), the rest goes
        through `load` one by one; images of equal decoded size then share one K5 launch per pass."""
        imsize = self.imsize if imsize is None else imsize
        decoded = [None] * len(items)
        jpg = [i for i, it in enumerate(items) if self.decode == "nvjpeg" and isinstance(it, (str, os.PathLike))
               and str(it).lower().endswith((".jpg", ".jpeg"))]
        if jpg:
            streams = []
            for i in jpg:
                with open(items[i], "rb") as f:
                    streams.append(f.read())
            for i, img in zip(jpg, _lib.jpeg_decode_batch(streams, self.device)):
                decoded[i] = img
        for i, it in enumerate(items):
            if decoded[i] is None:
                decoded[i] = self._decode(it)
        result = [None] * len(items)
        groups = {}
        for i, d in enumerate(decoded):
            groups.setdefault(tuple(d.shape), []).append(i)
        for shape, idxs in groups.items():
            if imsize is not None and len(idxs) > 1:
                batch = self.resize_batch([decoded[i] for i in idxs], imsize=imsize)
                for j, i in enumerate(idxs):
                    result[i] = batch[j]
            else:
                for i in idxs:
                    result[i] = self.resize(decoded[i], imsize=imsize)
        return result

    def load(self, item, bbx=None, imsize=None):
        if isinstance(item, np.ndarray):
            return self.crop_only(self._decode(item), bbx)
        return self.resize(self._decode(item), imsize=imsize, bbx=bbx)

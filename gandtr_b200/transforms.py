"""Image transforms of the retrieval path, host-side mirror of mdir/components/data/transform/.

Same registry, same string grammar and the same `repr` as the reference
(`TRANSFORMS`, `initialize_transforms(augmentations, mean_std)`, mdir/components/data/transform/__init__.py:3-45;
`GenericTransform.__repr__`, core_transforms.py:17-18), but the chain the hub models use,

    pil2np | apply_clahe:<clip>[:<grid>[:lab]] | totensor | normalize          (mdir/hub/embedding.yml:14)

is executed as ONE fused CUDA pass (K1, gdt_clahe_u8): uint8 HWC image in, normalised float32 CHW tensor out,
bit-identical to the reference's OpenCV path. The stage objects only carry parameters; `Compose` plans the chain once
and refuses chains that fall outside the hot path instead of silently running something else.

Because the work happens on the GPU the transform must be called from the main process (not inside DataLoader
workers); it returns a CUDA tensor, which the reference's `tensor.to(device)` accepts unchanged.
"""
import numpy as np
import torch

from . import _lib

__all__ = ["TRANSFORMS", "initialize_transforms", "Compose", "Pil2Numpy", "ApplyClahe", "ToTensor", "Normalize"]


class GenericTransform(object):
    def __init__(self, params=None):
        self.params = params or {}

    def __repr__(self):
        return self.__class__.__name__ + "(%s)" % ", ".join("%s=%s" % (x, str(y)) for x, y in self.params.items())


class Pil2Numpy(GenericTransform):
    """Convert pil image to numpy array with values between 0 and 1 (core_transforms.py:76-100). Fused: the uint8
    pixels travel to the device as they are, the /255.0 happens inside K1."""


class ApplyClahe(GenericTransform):
    """Convert input images to given colorspace and apply clahe to its lightness channel
    (photometric_transforms.py:28-36)."""

    def __init__(self, clip_limit=4, grid_size=8, colorspace="lab"):
        super().__init__({"clip_limit": float(clip_limit), "grid_size": int(grid_size), "colorspace": colorspace})
        if str(colorspace).lower() != "lab":
            raise NotImplementedError("gandtr_b200 implements apply_clahe for the 'lab' colorspace only (the hub "
                                      "models' setting, embedding.yml:14); got %r" % (colorspace,))


class ToTensor(GenericTransform):
    """HWC float image -> CHW tensor (core_transforms.py:35-44). Fused into K1's planar stores."""

    def __repr__(self):
        return self.__class__.__name__ + "()"


class Normalize(GenericTransform):
    """(x - mean) / std per channel (core_transforms.py:47-70)."""

    def __init__(self, mean, std, strict_shape=True):
        strict_shape = bool(strict_shape) if not isinstance(strict_shape, str) or strict_shape.lower() != "false" else False
        super().__init__({"mean": mean, "std": std, "strict_shape": strict_shape})
        assert len(mean) == len(std)


TRANSFORMS = {
    "totensor": ToTensor,
    "normalize": Normalize,
    "pil2np": Pil2Numpy,
    "apply_clahe": ApplyClahe,
}


def _as_u8_hwc(pic):
    """PIL image / uint8 ndarray -> contiguous uint8 HWC RGB ndarray (what np.asarray(pic.convert('RGB')) gives,
    core_transforms.py:96)."""
    if hasattr(pic, "convert"):
        pic = np.asarray(pic.convert("RGB"))
    if not isinstance(pic, np.ndarray):
        raise ValueError("Unsupported type '%s'" % type(pic))
    if pic.dtype != np.uint8 or pic.ndim != 3 or pic.shape[2] != 3:
        raise ValueError("the fused CLAHE transform takes 8-bit RGB images (H x W x 3 uint8), got %s %s"
                         % (pic.dtype, pic.shape))
    return np.ascontiguousarray(pic)


class Compose(object):
    """Sequential transform chain with the reference's calling convention (varargs in; a single result is unwrapped,
    core_transforms.py:26-32) that executes as one K1 launch per image -- or per same-size batch via `batch()`."""

    def __init__(self, transforms, device=None):
        self.transforms = list(transforms)
        self.device = torch.device(device) if device is not None else None
        kinds = [type(t) for t in self.transforms]
        if kinds == [Pil2Numpy, ApplyClahe, ToTensor, Normalize]:
            clahe, norm = self.transforms[1], self.transforms[3]
            self._clip, self._grid = clahe.params["clip_limit"], clahe.params["grid_size"]
        elif kinds == [Pil2Numpy, ToTensor, Normalize]:
            norm = self.transforms[2]
            self._clip = None
        else:
            raise NotImplementedError(
                "gandtr_b200 executes the retrieval chains 'pil2np | apply_clahe:... | totensor | normalize' and "
                "'pil2np | totensor | normalize'; got [%s]" % ", ".join(k.__name__ for k in kinds))
        self._mean = [float(x) for x in norm.params["mean"]]
        self._std = [float(x) for x in norm.params["std"]]
        if len(self._mean) != 3:
            raise NotImplementedError("3-channel RGB mean/std expected")

    def _device(self):
        if self.device is not None:
            return self.device
        if not torch.cuda.is_available():
            raise _lib.GdtError("gandtr_b200 transforms need a CUDA device (there is no CPU path)")
        return torch.device("cuda", torch.cuda.current_device())

    def batch(self, images_u8, out=None):
        """[n, h, w, 3] uint8 tensor (host, ideally pinned, or device) -> [n, 3, h, w] float32 CUDA tensor."""
        dev = self._device()
        x = images_u8 if images_u8.is_cuda else images_u8.to(dev, non_blocking=True)
        if self._clip is None:
            mean = torch.tensor(self._mean, dtype=torch.float32, device=dev).view(1, 3, 1, 1)
            std = torch.tensor(self._std, dtype=torch.float32, device=dev).view(1, 3, 1, 1)
            return (x.permute(0, 3, 1, 2).to(torch.float32) / 255.0).sub_(mean).div_(std)
        return _lib.clahe_u8(x.contiguous(), self._mean, self._std, clip_limit=self._clip, grid=self._grid, out=out)

    def __call__(self, *pics):
        outs = [self.batch(torch.from_numpy(_as_u8_hwc(p))[None])[0] for p in pics]
        if len(outs) == 1:
            return outs[0]
        return outs

    def __repr__(self):
        # torchvision.transforms.Compose.__repr__, which the reference inherits (README.md:133-138)
        format_string = self.__class__.__name__ + "("
        for t in self.transforms:
            format_string += "\n"
            format_string += "    {0}".format(t)
        format_string += "\n)"
        return format_string


def initialize_transforms(augmentations, mean_std, device=None):
    """mdir/components/data/transform/__init__.py:36-45 -- same parsing of 'a | b:arg1:arg2 | ...'."""
    trans = []
    for aug in [x.strip() for x in augmentations.split("|") if x.strip()]:
        tname, *args = aug.split(":", 1)
        args = args[0].split(":") if args else []
        if tname not in TRANSFORMS:
            raise NotImplementedError("transform '%s' is outside the retrieval hot path implemented by gandtr_b200 "
                                      "(available: %s)" % (tname, ", ".join(sorted(TRANSFORMS))))
        if "normalize" in aug:
            trans.append(TRANSFORMS[tname](*(list(mean_std) + args)))
        else:
            trans.append(TRANSFORMS[tname](*args))
    return Compose(trans, device=device)

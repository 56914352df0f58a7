"""Descriptor network, inference wrappers and the `SingleNetwork` facade -- host-side mirror of the reference objects
that sit on the retrieval hot path, with the arithmetic after the conv backbone executed by K2 (gdt_gem_whiten).

Reference objects mirrored (same names, constructor arguments, state_dict keys and return shapes):
  GeM, L2N                    mdir/external/cirtorch/layers/pooling.py:36-47, layers/normalization.py:10-20
  ImageRetrievalNet           mdir/external/cirtorch/networks/imageretrievalnet.py:90-143 (CirRetrievalNet: cirnet.py:8-45)
  init_network / init_cirnet  imageretrievalnet.py:146-309, mdir/components/model/network/cirnet.py:48-65
  Wrapper, Compose, CirMultiscaleAggregation, FakeBatch, CirFakeTupleBatch, CirtorchWhiten, ClahePost,
  WRAPPERS_LABELS, initialize_wrappers           mdir/components/data/wrapper.py:15-65,200-348,367-396
  SingleNetwork                                  mdir/learning/network.py:100-141

The VGG16 / ResNet-101 conv backbones stay stock torchvision modules (BASELINE.json north_star). What changes:
  * `ImageRetrievalNet.forward` = stock `features` -> ONE K2 call (GeM + L2N) instead of ~8 ATen launches;
  * `SingleNetwork.forward` recognises the hub models' eval wrapper stack {cirwhiten, cirmultiscale[, cirfaketuplebatch]}
    and runs per-scale backbones followed by ONE fused K2 call (GeM + L2N per scale, generalised-mean aggregation,
    renormalisation, centring, whitening projection, final L2N) -- no `.item()` sync, no per-scale launches;
  * the wrappers remain individually usable (`preprocess` / `postprocess` protocol) and route to the same kernels.
There is no CPU path: modules raise on CPU tensors.
"""
import copy
import hashlib
import json
import pickle
import random
import re
from collections import namedtuple
from typing import Any, NamedTuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .transforms import initialize_transforms

__all__ = ["GeM", "L2N", "ImageRetrievalNet", "init_network", "init_cirnet", "Wrapper", "Compose",
           "CirMultiscaleAggregation", "FakeBatch", "CirFakeTupleBatch", "CirtorchWhiten", "ClahePost",
           "RandomPassThrough", "CirRatioPassThrough", "MeanStdPost", "MeanStdPre", "MetadataTensor", "as_tensor",
           "as_metadata_tensor", "to_device", "WRAPPERS_LABELS", "initialize_wrappers", "SingleNetwork", "SequentialNetwork",
           "CirSequentialNetwork", "NETWORKS", "initialize_network", "OUTPUT_DIM"]

OUTPUT_DIM = {"alexnet": 256, "vgg11": 512, "vgg13": 512, "vgg16": 512, "vgg19": 512, "resnet18": 512, "resnet34": 512,
              "resnet50": 2048, "resnet101": 2048, "resnet152": 2048}


# ---------------------------------------------------------------------------------------------------- layers

class GeM(nn.Module):
    """Generalised-mean pooling; `p` is the learned 1-element parameter `pool.p` of reference checkpoints."""

    def __init__(self, p=3, eps=1e-6):
        super().__init__()
        self.p = nn.Parameter(torch.ones(1) * p)
        self.eps = eps

    def forward(self, x):
        """[N,C,h,w] -> [N,C,1,1] (LF.gem, layers/functional.py:21-22)."""
        return _lib.gem_pool(x.contiguous(), self.p.detach(), self.eps).view(x.shape[0], x.shape[1], 1, 1)

    def __repr__(self):
        return self.__class__.__name__ + "(" + "p=" + "{:.4f}".format(self.p.data.tolist()[0]) + ", " + "eps=" + str(self.eps) + ")"


class L2N(nn.Module):
    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        """x / (||x||_2 over dim 1 + eps) (LF.l2n, layers/functional.py:130-131) for [N,C] or [N,C,1,1]."""
        if x.dim() == 4 and x.shape[2] == 1 and x.shape[3] == 1 or x.dim() == 2:
            return _lib.l2n_rows(x.reshape(x.shape[0], x.shape[1]).contiguous(), self.eps).view(x.shape)
        raise NotImplementedError("L2N on spatial maps is outside the retrieval hot path of gandtr_b200")

    def __repr__(self):
        return self.__class__.__name__ + "(" + "eps=" + str(self.eps) + ")"


POOLING = {"gem": GeM}


class ImageRetrievalNet(nn.Module):
    """features (stock torchvision convs) -> GeM -> L2N, returned as D x N like the reference."""

    def __init__(self, features, lwhiten, pool, whiten, meta):
        super().__init__()
        if lwhiten is not None or whiten is not None:
            raise NotImplementedError("local / end-to-end whitening layers are outside the hot path (hub models use "
                                      "local_whitening=False, whitening=False: mdir/hub/embedding.yml:6-10)")
        self.features = features if isinstance(features, nn.Sequential) else nn.Sequential(*features)
        self.lwhiten = None
        self.pool = pool
        self.whiten = None
        self.norm = L2N()
        self.meta = meta

    def feature_map(self, x):
        return self.features(x).contiguous()

    def descriptors(self, fmaps, aggregate=False, msp_is_p=False, P=None, m=None, dim=None, P_split=None):
        """One K2 call on per-scale feature maps -> [N, dim]."""
        return _lib.gem_whiten(fmaps, self.pool.p.detach(), eps=self.pool.eps, aggregate=aggregate, msp_is_p=msp_is_p,
                               P=P, m=m, dim=dim, P_split=P_split)

    def forward(self, x):
        o = self.descriptors([self.feature_map(x)])         # norm(pool(o)).squeeze(-1).squeeze(-1), imageretrievalnet.py:116
        return o.permute(1, 0)                               # D x N (imageretrievalnet.py:123)

    # CirRetrievalNet (cirnet.py:8-45)
    def parameter_groups(self, optimizer_opts):
        return [{"params": self.features.parameters()},
                {"params": self.pool.parameters(), "lr": optimizer_opts["lr"] * 10, "weight_decay": 0}]

    @staticmethod
    def _set_batchnorm_eval(mod):
        if mod.__class__.__name__.find("BatchNorm") != -1:
            mod.eval()

    def train(self, mode=True):
        res = super().train(mode)
        if mode:
            self.apply(ImageRetrievalNet._set_batchnorm_eval)
        return res

    def meta_repr(self):
        tmpstr = "  (" + "meta" + "): dict( \n"
        for k, label in (("architecture", "architecture"), ("local_whitening", "local_whitening"), ("pooling", "pooling"),
                         ("regional", "regional"), ("whitening", "whitening"), ("out_channels", "outputdim")):
            tmpstr += "     {}: {}\n".format(label, self.meta[k])
        if "mean" in self.meta and "std" in self.meta:
            tmpstr += "     mean: {}\n".format(self.meta["mean"])
            tmpstr += "     std: {}\n".format(self.meta["std"])
        return tmpstr + "  )\n"

    def __repr__(self):
        return super().__repr__()[:-1] + self.meta_repr() + ")"


CirRetrievalNet = ImageRetrievalNet


def init_network(params):
    """imageretrievalnet.py:146-309 for the configurations on the hot path: torchvision backbone with random-init or
    checkpoint-loaded weights (no downloads), GeM pooling, no local / regional / end-to-end whitening."""
    import torchvision
    architecture = params.get("architecture", "resnet101")
    pooling = params.get("pooling", "gem")
    if params.get("local_whitening", False) or params.get("regional", False) or params.get("whitening", False):
        raise NotImplementedError("local_whitening / regional / whitening layers are outside the hot path")
    if pooling not in POOLING:
        raise NotImplementedError("pooling '%s' is outside the hot path (available: gem)" % (pooling,))
    if params.get("pretrained", False):
        raise NotImplementedError("pretrained=True needs network access (ImageNet / retrieval-SfM weights are HTTP downloads: "
                                  "imageretrievalnet.py:21-31); load a checkpoint with load_state_dict instead")
    net_in = getattr(torchvision.models, architecture)(weights=None)
    if architecture.startswith("alexnet") or architecture.startswith("vgg"):
        features = list(net_in.features.children())[:-1]
    elif architecture.startswith("resnet") or architecture.startswith("resnext"):
        features = list(net_in.children())[:-2]
    else:
        raise ValueError("Unsupported or unknown architecture: {}!".format(architecture))
    last_convs = [x for x in list(features[-2].modules()) + list(features[-1].modules()) if isinstance(x, nn.Conv2d)]
    dim = last_convs[-1].out_channels
    meta = {"architecture": architecture, "local_whitening": False, "pooling": pooling, "regional": False,
            "whitening": False, "mean": params.get("mean", [0.485, 0.456, 0.406]),
            "std": params.get("std", [0.229, 0.224, 0.225]), "outputdim": dim, "out_channels": dim}
    return ImageRetrievalNet(features, None, POOLING[pooling](), None, meta)


def init_cirnet(**params):
    """cirnet.py:48-65."""
    for key in ["local_whitening", "pooling", "regional", "whitening", "pretrained"]:
        if key not in params:
            raise ValueError("Key '%s' not in params" % key)
    params = dict(params)
    params["mean"] = [0.485, 0.456, 0.406]
    params["std"] = [0.229, 0.224, 0.225]
    params["architecture"] = params.pop("cir_architecture")
    net = init_network(params)
    net.meta["in_channels"] = 3
    net.meta["out_channels"] = net.meta["outputdim"]
    return net


# ---------------------------------------------------------------------------------------------------- wrappers

class MetadataTensor(NamedTuple):
    """(tensor, metadata) pair the reference's datasets hand to `network(x)` (mdir/tools/tensors.py:37-64): a NamedTuple
    for `default_collate`, with `.to` / `.unsqueeze_` propagated and every other attribute read from the tensor."""

    tensor: Any
    metadata: Any

    def __repr__(self):
        return "Data:\n%s\nMetadata:\n%s" % (self.tensor, self.metadata)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        metadata, = (a.metadata for a in args if isinstance(a, cls))
        args = [a.tensor if isinstance(a, cls) else a for a in args]
        return MetadataTensor(func(*args, **(kwargs or {})), metadata)

    def __getattr__(self, name):
        attr = getattr(self.tensor, name)
        if name not in ["to", "unsqueeze_"]:
            return attr

        def func(*args, **kwargs):
            return self.__class__(attr(*args, **kwargs), self.metadata)
        return func


def as_metadata_tensor(tensor, metadata):
    """tools/tensors.py:67-72."""
    if isinstance(tensor, MetadataTensor):
        assert not set(tensor.metadata.keys()) & set(metadata.keys())
        tensor.metadata.update(dict(metadata))
        return tensor
    return MetadataTensor(torch.as_tensor(tensor), dict(metadata))


def as_tensor(tensor):
    """Strip the metadata from tensors nested in lists / tuples / dicts (tools/tensors.py:74-85)."""
    if tensor is None:
        return None
    if isinstance(tensor, MetadataTensor):
        return tensor.tensor
    if isinstance(tensor, list):
        return [as_tensor(x) for x in tensor]
    if isinstance(tensor, tuple):
        return tuple(as_tensor(x) for x in tensor)
    if isinstance(tensor, dict):
        return {k: as_tensor(v) for k, v in tensor.items()}
    return torch.as_tensor(tensor)


def to_device(tensor, device):
    """Move tensors nested in lists / tuples / dicts to `device`, keeping the structure (tools/tensors.py:8-20)."""
    if isinstance(tensor, MetadataTensor):                      # before the tuple test: a NamedTuple is a tuple
        return MetadataTensor(tensor.tensor.to(device), tensor.metadata)
    if hasattr(tensor, "to"):
        return tensor.to(device)
    if isinstance(tensor, list):
        return [to_device(x, device) for x in tensor]
    if isinstance(tensor, tuple):
        return tuple(to_device(x, device) for x in tensor)
    if isinstance(tensor, dict):
        return {k: to_device(v, device) for k, v in tensor.items()}
    if tensor is None:
        return None
    return tensor.to(device)


_to_device = to_device


def _plain(x):
    """The torch.Tensor inside a MetadataTensor (kernels take raw device pointers), anything else unchanged."""
    return x.tensor if isinstance(x, MetadataTensor) else x


class Compose(object):
    """Wrappers applied in order for preprocess and in reverse for postprocess (wrapper.py:15-49)."""

    def __init__(self, wrappers, device):
        self.wrappers = wrappers
        self.device = device

    def __call__(self, tensor, inference, outputmodel=None, tensor_params=None):
        tensor_params = {} if tensor_params is None else tensor_params
        if not self.wrappers:
            if isinstance(tensor, torch.Tensor):
                tensor = tensor.to(self.device)
            return inference(tensor, **tensor_params)
        if outputmodel is None:
            outputmodel = inference
        metadata = []
        for wrapper in self.wrappers:
            tensor, meta = wrapper.preprocess(tensor, outputmodel)
            metadata.append(meta)
        tensor = inference(_to_device(tensor, self.device), **tensor_params)
        for wrapper, meta in reversed(list(zip(self.wrappers, metadata))):
            tensor = wrapper.postprocess(tensor, outputmodel, meta)
        return tensor

    def __repr__(self):
        nice = "\n" + "".join("    %s\n" % x for x in self.wrappers) if self.wrappers else ""
        return "%s([%s])" % (self.__class__.__name__, nice)


class Wrapper(object):
    def __init__(self, device):
        pass

    def preprocess(self, tensor, _outputmodel):
        return tensor, None

    def postprocess(self, tensor, _outputmodel, _metadata):
        return tensor


class RandomPassThrough(Wrapper):
    """Lets the input through the wrapped network with the given probability, otherwise the input skips it
    (wrapper.py:97-117)."""

    def __init__(self, probability_through, device):
        super().__init__(device)
        self.probability = float(probability_through)

    def preprocess(self, tensor, outputmodel):
        if isinstance(tensor, list):
            out = tuple(zip(*[self.preprocess(t, outputmodel) for t in tensor]))
            return list(out[0]), list(out[1])
        return (tensor, None) if random.random() < self.probability else (None, tensor)

    def postprocess(self, tensor, outputmodel, tensor_skipping):
        if isinstance(tensor, list):
            return [self.postprocess(t, outputmodel, s) for (t, s) in zip(tensor, tensor_skipping)]
        return tensor if tensor_skipping is None else as_tensor(tensor_skipping)

    def __repr__(self):
        return "%s(probability=%s)" % (self.__class__.__name__, self.probability)


class CirRatioPassThrough(RandomPassThrough):
    """Deterministic variant for the retrieval training tuples: an image goes through the generator iff its
    `image_label` matches and the md5 of its name falls below the ratio (wrapper.py:120-146)."""

    def __init__(self, ratio_through, image_label, *, device):
        super().__init__(ratio_through, device)
        self.image_label = re.compile(image_label)

    def preprocess(self, tensor, outputmodel):
        if isinstance(tensor, list):
            acc = [self.preprocess(x, outputmodel) for x in tensor]
            return tuple(list(x) for x in zip(*acc))
        image_label = tensor.metadata["image_label"]            # must be present (sanity check of the reference)
        if isinstance(image_label, list) and len(image_label) == 1:
            image_label = image_label[0]
        if self.image_label.match(image_label) and self._passthrough(tensor.metadata["name"]):
            return tensor, None
        return None, tensor

    def _passthrough(self, name):
        if isinstance(name, list):
            name, = name
        digits = 4
        rand = int(hashlib.md5(name.encode("utf8")).hexdigest()[-digits:], 16) / (16 ** digits)
        return rand < self.probability

    def __repr__(self):
        return "%s(probability=%s, train_label=%s)" % (self.__class__.__name__, self.probability, self.image_label)


class MeanStdPost(Wrapper):
    """Adapts the normalisation of the network output: x*std_in + mean_in, then (. - mean_out) / std_out
    (wrapper.py:149-179). One elementwise kernel (gdt_meanstd_adapt) with the reference's four roundings."""

    def __init__(self, input_meanstd, output_meanstd, device):
        super().__init__(device)
        input_meanstd = json.loads(input_meanstd) if isinstance(input_meanstd, str) else input_meanstd
        output_meanstd = json.loads(output_meanstd) if isinstance(output_meanstd, str) else output_meanstd
        if any(x == 0 for x in input_meanstd[1]) or any(x == 0 for x in output_meanstd[1]):
            raise ValueError("Some std element is zero, leading to zero division.")
        self.input_meanstd = [self.mean2tensor(x, device) for x in input_meanstd]
        self.output_meanstd = [self.mean2tensor(x, device) for x in output_meanstd]
        self.device = device
        self._host = ([float(v) for v in input_meanstd[0]], [float(v) for v in input_meanstd[1]],
                      [float(v) for v in output_meanstd[0]], [float(v) for v in output_meanstd[1]])

    @staticmethod
    def mean2tensor(mean, device):
        mean = torch.as_tensor(mean, device=device)
        if mean.ndim == 1:
            mean = mean[:, None, None]
        return mean

    def postprocess(self, tensor, outputmodel, meta):
        if isinstance(tensor, list):
            return [self.postprocess(x, outputmodel, meta) for x in tensor]
        return self._adapt(tensor)

    def _adapt(self, tensor):
        if tensor is None:
            return None
        t = _plain(tensor)
        out = _lib.meanstd_adapt(t.detach().to(self.device).contiguous(), *self._host)
        return MetadataTensor(out, tensor.metadata) if isinstance(tensor, MetadataTensor) else out

    def __repr__(self):
        return "%s(input_meanstd=%s,output_meanstd=%s)" % (self.__class__.__name__, self.input_meanstd, self.output_meanstd)


class MeanStdPre(MeanStdPost):
    """The same adaptation applied to the network input (wrapper.py:182-194)."""

    def preprocess(self, tensor, _outputmodel):
        if isinstance(tensor, list):
            return [self.preprocess(x, _outputmodel) for x in tensor]
        return self._adapt(tensor), None

    def postprocess(self, tensor, outputmodel, meta):
        return tensor


class CirMultiscaleAggregation(Wrapper):
    """Downscale each image to defined scales and aggregate resulting descriptors (wrapper.py:200-263)."""

    def __init__(self, scales, device):
        super().__init__(device)
        if isinstance(scales, str):
            scales = {"True": True, "False": False, "ms": True, "ss": False,
                      "sms5": [1, 1. / np.sqrt(2), np.sqrt(2), 1. / 2, 2],
                      "sms": [1, 1. / np.sqrt(2), np.sqrt(2)]}[scales]
        if isinstance(scales, bool):
            scales = [1, 1. / np.sqrt(2), 1. / 2] if scales else [1]
        self.scales = scales

    @staticmethod
    def interpolate(x, scale):
        """wrapper.py:212-233: stock ATen bilinear resize; a MetadataTensor keeps its metadata (`wrap_metadata`)."""
        if isinstance(x, MetadataTensor):
            return MetadataTensor(F.interpolate(x.tensor, scale_factor=scale, mode="bilinear", align_corners=False), x.metadata)
        return F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=False)

    def preprocess(self, tensor, _outputmodel):
        if len(self.scales) == 1:
            return tensor if isinstance(tensor, list) else [tensor], isinstance(tensor, list)
        if isinstance(tensor, list):
            return [self.interpolate(single, scale) for single in tensor for scale in self.scales], True
        return [self.interpolate(tensor, scale) for scale in self.scales], False

    @staticmethod
    def aggregate_tensor(tensor, nscales, outputdim, msp):
        """v = (mean_s d_s^msp)^(1/msp); v /= ||v|| (wrapper.py:235-245) via gdt_desc_post; `msp` may be a float or a
        1-element device tensor (no host sync in the latter case)."""
        assert len(tensor) == nscales, "%s != %s" % (len(tensor), nscales)
        descs = [t.reshape(outputdim, -1).t().contiguous() for t in tensor]            # D x N -> N x D
        out = _lib.desc_post(descs, msp)
        return out.squeeze(0) if out.shape[0] == 1 else out.t()

    @staticmethod
    def _msp(scales, outputmodel):
        if len(scales) > 1 and outputmodel.meta.get("pooling", None) == "gem" \
                and not outputmodel.meta["regional"] and not outputmodel.meta["whitening"]:
            return outputmodel.pool.p.detach()      # the reference calls .item() here (wrapper.py:251): a device sync we skip
        return 1.0

    def postprocess(self, tensor, outputmodel, waslist):
        msp = self._msp(self.scales, outputmodel)
        dim = outputmodel.meta["out_channels"]
        if not waslist:
            return self.aggregate_tensor(tensor, len(self.scales), dim, msp)
        assert len(tensor) % len(self.scales) == 0, "%s %% %s != 0" % (len(tensor), len(self.scales))
        return [self.aggregate_tensor(tensor[i:i + len(self.scales)], len(self.scales), dim, msp)
                for i in range(0, len(tensor), len(self.scales))]

    def __repr__(self):
        return "%s(scales=%s)" % (self.__class__.__name__, self.scales)


class FakeBatch(Wrapper):
    """Mimic batch behaviour by accumulating the result across multiple images (wrapper.py:266-280)."""

    def postprocess(self, tensor, outputmodel, _meta):
        if not isinstance(tensor, list) or not isinstance(tensor[0], torch.Tensor):
            return tensor
        return torch.stack([vec.reshape(-1) for vec in tensor], dim=1)      # out_channels x len(tensor)

    def __repr__(self):
        return "%s()" % self.__class__.__name__


class CirFakeTupleBatch(FakeBatch):
    """wrapper.py:283-305."""

    @classmethod
    def unsqueeze(cls, tensor):
        if isinstance(tensor, list):
            return [cls.unsqueeze(x) for x in tensor]
        elif len(tensor.shape) == 3:
            return tensor.unsqueeze_(0)
        elif len(tensor.shape) == 4:
            return tensor
        raise ValueError("Unsupported tensor dimensionality %s" % len(tensor.shape))

    def preprocess(self, tensor, _outputmodel):
        if not isinstance(tensor, list) or not isinstance(tensor[0], list):
            return tensor, False
        acc = []
        meta = len(tensor[0])
        for tpl in tensor:
            assert meta == len(tpl)
            acc += tpl
        return acc, meta


def _load_whitening(whitening):
    """{'m': D x 1, 'P': D x D} pickle (mdir/stages/whiten.py:75) or an in-memory dict."""
    if isinstance(whitening, dict):
        return whitening
    if isinstance(whitening, str) and whitening.startswith(("http://", "https://")):
        raise NotImplementedError("whitening %s is an HTTP download (mdir/hub/model.py:60-61); pass a local .pkl path or a dict"
                                  % whitening)
    with open(whitening, "rb") as handle:
        return pickle.load(handle)


class CirtorchWhiten(Wrapper):
    """Whiten vectors with possible dimensionality reduction (wrapper.py:308-322)."""

    def __init__(self, whitening, dimensions, device):
        super().__init__(device)
        whitening = _load_whitening(whitening)
        self.P = torch.tensor(np.asarray(whitening["P"]), dtype=torch.float32, device=device).contiguous()
        self.m = torch.tensor(np.asarray(whitening["m"]), dtype=torch.float32, device=device).contiguous()
        self.dimensions = dimensions or self.P.shape[0]
        # the projection is a learned constant: split it once into TF32 hi / lo halves for the tcgen05 kernel
        self.P_split = _lib.whiten_prepare(self.P[:self.dimensions].contiguous()) if self.P.is_cuda else None

    def postprocess(self, tensor, _outputmodel, _meta):
        """[D] (one image) -> [dim];  [D, N] -> [dim, N]."""
        single = tensor.dim() == 1
        v = tensor.reshape(self.P.shape[1], -1).t().contiguous()
        out = _lib.desc_post([v], None, P=self.P, m=self.m.reshape(-1), dim=self.dimensions, P_split=self.P_split)
        return out.squeeze(0) if single else out.t()

    def __repr__(self):
        return "%s(dimensions=%s)" % (self.__class__.__name__, self.dimensions)


class ClahePost(Wrapper):
    """CLAHE applied to a normalised CHW device tensor (wrapper.py:325-348) without leaving the GPU (gdt_clahe_f32)."""

    def __init__(self, meanstd, clip_limit=4, grid_size=8, colorspace="lab", *, device):
        super().__init__(device)
        meanstd = json.loads(meanstd) if isinstance(meanstd, str) else meanstd
        self.mean, self.std = [float(x) for x in meanstd[0]], [float(x) for x in meanstd[1]]
        self.clip_limit, self.grid_size = float(clip_limit), int(grid_size)
        self.device = device
        if str(colorspace).lower() != "lab":
            raise NotImplementedError("ClahePost: only the 'lab' colorspace is implemented")

    def postprocess(self, tensor, outputmodel, meta):
        if tensor is None:
            return tensor
        if isinstance(tensor, list):
            return [self.postprocess(x, outputmodel, meta) for x in tensor]
        if len(tensor.shape) == 4:
            # images that skipped the generator (cir_ratio_pass_through) may still be host tensors: CLAHE runs on the device
            return _lib.clahe_f32(_plain(tensor).detach().to(self.device).contiguous(), self.mean, self.std, self.mean,
                                  self.std, clip_limit=self.clip_limit, grid=self.grid_size)
        if len(tensor.shape) == 3:
            return self.postprocess(tensor.unsqueeze(0), outputmodel, meta)[0]
        raise ValueError("Unsupported tensor dims: %s" % len(tensor.shape))

    def __repr__(self):
        return "%s(clip_limit=%s, grid_size=%s)" % (self.__class__.__name__, self.clip_limit, self.grid_size)


WRAPPERS_LABELS = {
    "random_pass_through": RandomPassThrough,
    "cir_ratio_pass_through": CirRatioPassThrough,
    "meanstd_post": MeanStdPost,
    "meanstd_pre": MeanStdPre,
    "cirmultiscale": CirMultiscaleAggregation,
    "fakebatch": FakeBatch,
    "cirfaketuplebatch": CirFakeTupleBatch,
    "cirwhiten": CirtorchWhiten,
    "clahepost": ClahePost,
}


def initialize_wrappers(net_wrappers, device):
    """wrapper.py:384-396: '' / None, 'label:arg,label' strings, or {'<order>_<label>': kwargs} dicts."""
    if net_wrappers is None:
        wraps = []
    elif isinstance(net_wrappers, str):
        wraps = []
        for wrap in [x.strip() for x in _split_outside_brackets(net_wrappers, ",") if x.strip()]:
            wname, *args = wrap.split(":")
            wraps.append(_wrapper_class(wname)(*args, device=device))
    else:
        wraps = [_wrapper_class(x.split("_", 1)[1])(**net_wrappers[x], device=device) for x in sorted(net_wrappers)]
    return Compose(wraps, device)


def _split_outside_brackets(seq, sep):
    """Split on `sep` outside (), [], {} -- wrapper arguments such as [[0.5,0.5,0.5],[0.5,0.5,0.5]] contain commas
    (mdir/tools/utils.py:95-112 `splitp` with check_valid_pairs)."""
    parts, depth, closing = [""], [], {"(": ")", "[": "]", "{": "}"}
    for ch in seq:
        if ch == sep and not depth:
            parts.append("")
            continue
        if ch in closing:
            depth.append(closing[ch])
        elif depth and ch == depth[-1]:
            depth.pop()
        parts[-1] += ch
    assert not depth, 'Invalid seq "%s": unbalanced brackets' % seq
    return parts


def _wrapper_class(name):
    if name not in WRAPPERS_LABELS:
        raise NotImplementedError("wrapper '%s' is outside the retrieval hot path implemented by gandtr_b200 (available: %s)"
                                  % (name, ", ".join(sorted(WRAPPERS_LABELS))))
    return WRAPPERS_LABELS[name]


# ---------------------------------------------------------------------------------------------------- network facade

class SingleNetwork(object):
    """mdir/learning/network.py:100-141: model + stage-dependent wrappers; `__call__` is inference."""

    TRAIN = "train"
    EVAL = "eval"
    NetworkParams = namedtuple("NetworkParams", ["model", "runtime"])

    def __init__(self, model, network_params, device, frozen):
        self.meta = model.meta if model.meta else {}
        self.network_params = network_params
        runtime = network_params.runtime
        assert not runtime.keys() - {"data", "wrappers", "frozen", "model"}, runtime.keys() - {"data", "wrappers", "frozen", "model"}
        assert not runtime.get("data", {}).keys() - {"mean_std", "transforms", "augmentations"}
        wrappers = runtime.get("wrappers", "")
        if isinstance(wrappers, dict) and wrappers.keys() == {"train", "eval"}:
            self.wrappers = {x: initialize_wrappers(wrappers[x], device) for x in wrappers}
        else:
            self.wrappers = {x: initialize_wrappers(wrappers, device) for x in ["train", "eval"]}
        self.frozen = runtime.get("frozen", False) or frozen
        self.model = model.to(device)
        self.device = torch.device(device)
        self.stage = None
        if self.frozen:
            self.eval()

    def train(self):
        if not self.frozen:
            self.model.train()
            self.stage = self.TRAIN
        return self

    def eval(self):
        self.model.eval()
        self.stage = self.EVAL
        return self

    def __call__(self, image):
        return self.forward(image)

    def _fused_plan(self):
        """(whiten wrapper | None, multiscale wrapper) when the active stack is the hub models' eval stack."""
        wraps = self.wrappers[self.stage].wrappers
        kinds = [type(w) for w in wraps]
        if kinds and kinds[-1] is CirFakeTupleBatch:
            kinds, wraps = kinds[:-1], wraps[:-1]
        if kinds == [CirtorchWhiten, CirMultiscaleAggregation]:
            return wraps[0], wraps[1]
        if kinds == [CirMultiscaleAggregation]:
            return None, wraps[0]
        return None

    def forward(self, image, **params):
        """`network(x)`: x is a tensor, a MetadataTensor (tools/tensors.py:37), or lists / tuples of them -- whatever the
        active wrappers' `preprocess` accept (network.py:133-134 of the reference)."""
        plan = None if params else self._fused_plan()
        if plan is not None:
            # the fused stack reads no metadata: MetadataTensors are unwrapped the way forward_batch's as_tensor does
            bare = as_tensor(image) if isinstance(image, (MetadataTensor, list)) else image
            if isinstance(bare, torch.Tensor) and bare.dim() == 4:
                return self._forward_fused(bare, *plan)
            if isinstance(bare, list) and bare and all(isinstance(x, torch.Tensor) and x.dim() == 4 for x in bare):
                return [self._forward_fused(x, *plan) for x in bare]
        return self.wrappers[self.stage](image, self.forward_batch, outputmodel=self.model, tensor_params=params)

    def _forward_fused(self, image, whiten, ms):
        """[N,3,H,W] -> per-scale stock backbones -> ONE K2 launch group. Returns [dim] for N == 1 (as the reference's
        aggregate_tensor / whitening squeeze does) and [dim, N] otherwise."""
        x = image.to(self.device)
        model = self.model
        scales = ms.scales
        if len(scales) == 1:
            fmaps = [model.feature_map(x)]
        else:
            fmaps = [model.feature_map(ms.interpolate(x, s)) for s in scales]     # every scale, incl. 1 (wrapper.py:218-233)
        msp_is_p = not isinstance(ms._msp(scales, model), float)
        d = model.descriptors(fmaps, aggregate=True, msp_is_p=msp_is_p,
                              P=whiten.P if whiten is not None else None,
                              m=whiten.m.reshape(-1) if whiten is not None else None,
                              dim=whiten.dimensions if whiten is not None else None,
                              P_split=whiten.P_split if whiten is not None else None)
        return d.squeeze(0) if d.shape[0] == 1 else d.t()

    def forward_batch(self, images, **params):
        """network.py:136-141: metadata is dropped right before the model (`tensors.as_tensor`); dict / tuple inputs reach
        the model with their structure intact."""
        if images is None:
            return None
        if isinstance(images, list):
            return [self.model(as_tensor(x), **params) if x is not None else None for x in images]
        return self.model(as_tensor(images), **params)

    @classmethod
    def initialize(cls, params, device):
        """network.py:142-190 for parameter dicts without a remote checkpoint: {model, runtime, initialize[, path]}."""
        params = copy.deepcopy(params)
        path = params.pop("path", None)
        init = params.pop("initialize", None)
        if path:
            checkpoint = torch.load(path, map_location="cpu", weights_only=False)
            if "net" in checkpoint:
                checkpoint = checkpoint["net"]
            runtime = params.pop("runtime")
            if runtime == "load_from_checkpoint":
                runtime = checkpoint["network_params"]["runtime"]
            else:
                runtime = {x: y if y != "load_from_checkpoint" else checkpoint["network_params"]["runtime"][x]
                           for x, y in runtime.items()}
            network_params = cls.NetworkParams(checkpoint["network_params"]["model"], runtime)
            model = initialize_model(copy.deepcopy(network_params.model))
            model.load_state_dict(checkpoint["model_state"])
            params.pop("model", None)
        else:
            network_params = cls.NetworkParams(params.pop("model"), params.pop("runtime"))
            model = initialize_model(copy.deepcopy(network_params.model))
            if init and isinstance(init, str):
                model.load_state_dict(torch.load(init, map_location="cpu"))
            elif isinstance(init, dict) and init.get("weights") in ("normal_p2p", "kaiming_p2p"):
                # generator.yml:11-13 (mdir/components/model/weight_initialization.py:62-76)
                from .generator import init_weights_p2p
                if init.get("seed") is not None:
                    torch.manual_seed(init["seed"])
                kind = init["weights"].split("_")[0]
                model.apply(lambda mod: init_weights_p2p(mod, kind, 0.2))
            elif init:
                raise NotImplementedError("weight initialisation %s is outside the hot path" % (init,))
        params.pop("type", None)
        assert not params, params.keys()
        return cls(model, network_params, device=device, frozen=False)

    def state_dict(self):
        return {"net": {"type": self.__class__.__name__, "frozen": self.frozen,
                        "network_params": self.network_params._asdict(), "model_state": self.model.state_dict()}}


def initialize_model(model_params):
    """mdir/components/model/network/__init__.py: only the 'cirnet' architecture is on the hot path."""
    model_params = dict(model_params)
    arch = model_params.pop("architecture")
    if arch == "official_resnet_generator":                     # mdir/hub/generator.yml:3-10 (stock PyTorch, north_star)
        from .generator import ResnetGenerator
        return ResnetGenerator(**model_params)
    if arch != "cirnet":
        raise NotImplementedError("architecture '%s' is outside the retrieval hot path (available: cirnet, "
                                  "official_resnet_generator)" % arch)
    return init_cirnet(**model_params)


class SequentialNetwork(object):
    """Two networks run back to back, e.g. `augment,embed` = day-to-night generator -> descriptor network
    (mdir/learning/network.py:635-677; BASELINE config 5, iccv23/parameters/finetune.yml:5-32). The last network's
    wrappers become the sequence's own (they see the raw input and the final output); the first network keeps its
    wrappers (meanstd_post, clahepost, cir_ratio_pass_through for the generator), so everything between the two models
    stays on the device: generator -> K1' (gdt_clahe_f32) -> gdt_meanstd_adapt -> backbone -> K2."""

    NetworkParams = namedtuple("NetworkParams", ["runtime"])

    def __init__(self, networks, sequence, device, frozen, rearrange_wrappers=True):
        assert len(networks) == 2
        assert networks.keys() == set(sequence)
        self.networks, self.network_order = networks, list(sequence)
        first_net = networks[sequence[0]]
        self.last_net = networks[sequence[1]]
        self.model = self.last_net.model
        self.device = torch.device(device)
        self.frozen = frozen
        if rearrange_wrappers:
            self.wrappers = self.last_net.wrappers
            self.last_net.wrappers = {x: initialize_wrappers("", device) for x in ["train", "eval"]}
            self.network_params = self.NetworkParams({"wrappers": self.last_net.network_params.runtime.get("wrappers", ""),
                                                      "data": first_net.network_params.runtime["data"]})
        else:
            self.wrappers = {x: initialize_wrappers("", device) for x in ["train", "eval"]}
            self.network_params = self.NetworkParams({"wrappers": "", "data": first_net.network_params.runtime["data"]})
        assert first_net.meta["out_channels"] == self.last_net.meta["in_channels"]
        self.meta = {"in_channels": first_net.meta["in_channels"], "out_channels": self.last_net.meta["out_channels"]}
        self.stage = None
        if frozen:
            self.eval()

    def train(self):
        for net in self.networks.values():
            net.train()
        self.stage = SingleNetwork.TRAIN
        return self

    def eval(self):
        for net in self.networks.values():
            net.eval()
        self.stage = SingleNetwork.EVAL
        return self

    def __call__(self, image):
        return self.forward(image)

    def forward(self, image):
        return self.wrappers[self.stage](image, self.forward_batch, outputmodel=self.model)

    def forward_batch(self, images):
        if images is None:
            return None
        if isinstance(images, list):
            return [self._forward_all(x) for x in images]
        return self._forward_all(images)

    def _forward_all(self, image):
        for net in self.network_order:
            image = self.networks[net](image)
        return image

    @classmethod
    def initialize(cls, params, device):
        # already built networks may be passed in place of their parameter dicts (they are used, not copied)
        params = {k: (v if hasattr(v, "forward_batch") else copy.deepcopy(v)) for k, v in params.items()}
        sequence = params.pop("sequence").split(",")
        rearrange = params.pop("rearrange_wrappers") if "rearrange_wrappers" in params else True
        params.pop("type", None)
        networks = {x: (params[x] if hasattr(params[x], "forward_batch") else initialize_network(params[x], device))
                    for x in params}
        return cls(networks, sequence, device=device, frozen=False, rearrange_wrappers=rearrange)


class CirSequentialNetwork(SequentialNetwork):
    """network.py:750-756: the tuple list produced by `cirfaketuplebatch` goes through the sequence as ONE list, so the
    first network's wrappers see every image of the tuple (and its metadata) at once."""

    def forward_batch(self, images):
        if images is None:
            return None
        return self._forward_all(images)


NETWORKS = {"SingleNetwork": SingleNetwork, "SequentialNetwork": SequentialNetwork, "CirSequentialNetwork": CirSequentialNetwork}


def initialize_network(params, device):
    """mdir/learning/network.py `initialize_network`: {type, ...} -> network object."""
    params = dict(params)
    return NETWORKS[params.pop("type", "SingleNetwork")].initialize(params, device)


def attach_transform(network):
    """mdir/hub/model.py:38-42: `.transform` built from the network's runtime data parameters."""
    data_params = network.network_params.runtime["data"]
    if "augmentations" not in data_params:
        data_params["augmentations"] = data_params.pop("transforms")
    network.transform = initialize_transforms(data_params["augmentations"], data_params["mean_std"], device=network.device)
    return network

"""torch.hub entry points -- same names and keyword arguments as the reference (hubconf.py:1-4,
mdir/hub/model.py:52-154): gem_vgg16_cyclegan, gem_vgg16_hedngan, gem_resnet101_cyclegan, gem_resnet101_hedngan,
cyclegan, hedngan; each `f(pretrained=True, device=None)` returns an eval-mode network with `.transform`.

Differences forced by the environment, not by design: there is no CPU path (device defaults to the current CUDA device
and a CPU device is refused) and `pretrained=True` cannot download (`BASE_URL` files, model.py:5): point the
GANDTR_B200_WEIGHTS environment variable (or the `weights_dir` keyword) at a directory holding the reference's
`<name>.pth` / `<name>_lw.pkl` files to load them.
"""
import copy
import os

import torch

from . import network as N
from .generator import ResnetGenerator, init_weights_p2p

__all__ = ["gem_vgg16_cyclegan", "gem_vgg16_hedngan", "gem_resnet101_cyclegan", "gem_resnet101_hedngan", "cyclegan", "hedngan"]

MEAN_STD = [[0.485, 0.456, 0.406], [0.229, 0.224, 0.225]]

# mdir/hub/embedding.yml
EMBEDDING = {
    "initialized": {
        "type": "SingleNetwork",
        "model": {"architecture": "cirnet", "cir_architecture": None, "local_whitening": False, "pooling": "gem",
                  "pretrained": False, "regional": False, "whitening": False},
        "initialize": False,
        "runtime": {"data": {"transforms": "pil2np | apply_clahe:1.0 | totensor | normalize", "mean_std": MEAN_STD},
                    "wrappers": "cirfaketuplebatch"},
    },
    "pretrained": {
        "path": None,
        "runtime": {"data": "load_from_checkpoint",
                    "wrappers": {"train": None, "eval": {"0_cirwhiten": {"whitening": None, "dimensions": None},
                                                         "1_cirmultiscale": {"scales": True}}}},
    },
}


def _device(device):
    if not device:
        if not torch.cuda.is_available():
            raise RuntimeError("gandtr_b200 needs a CUDA device (B200); there is no CPU path")
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("gandtr_b200 has no CPU path: device must be a CUDA device, got %s" % device)
    return device


def _weights_file(weights_dir, name):
    weights_dir = weights_dir or os.environ.get("GANDTR_B200_WEIGHTS")
    if not weights_dir:
        raise NotImplementedError(
            "pretrained=True downloads %s from the authors' server (mdir/hub/model.py:5), which this environment cannot "
            "reach; set GANDTR_B200_WEIGHTS (or weights_dir=) to a directory with the file, or use pretrained=False" % name)
    path = os.path.join(weights_dir, name)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return path


def _embedding(arch, tag, pretrained, device, weights_dir):
    device = _device(device)
    if pretrained:
        params = copy.deepcopy(EMBEDDING["pretrained"])
        params["path"] = _weights_file(weights_dir, "%s_embed_%s.pth" % (tag, arch))
        params["runtime"]["wrappers"]["eval"]["0_cirwhiten"]["whitening"] = _weights_file(
            weights_dir, "%s_embed_%s_lw.pkl" % (tag, arch))
        state = torch.load(params["path"], map_location="cpu", weights_only=False)
        state = state["net"] if "net" in state else state
        runtime = params["runtime"]
        runtime = {k: (v if v != "load_from_checkpoint" else state["network_params"]["runtime"][k]) for k, v in runtime.items()}
        model_params = dict(state["network_params"]["model"])
        model_params["pretrained"] = False                     # model.py:31-33: no ImageNet download for cirnet
        model = N.initialize_model(model_params)
        model.load_state_dict(state["model_state"])
        net = N.SingleNetwork(model, N.SingleNetwork.NetworkParams(model_params, runtime), device, frozen=True)
    else:
        params = copy.deepcopy(EMBEDDING["initialized"])
        params["model"]["cir_architecture"] = arch
        params.pop("type")
        net = N.SingleNetwork.initialize(params, device)
    return N.attach_transform(net.eval())


def gem_vgg16_cyclegan(pretrained=True, device=None, weights_dir=None):
    """GeM global descriptor model with VGG16 backbone, CycleGAN query augmentation and CLAHE (model.py:52-66)."""
    return _embedding("vgg16", "cyclegan", pretrained, device, weights_dir)


def gem_vgg16_hedngan(pretrained=True, device=None, weights_dir=None):
    """GeM VGG16, HED^N-GAN augmentation and CLAHE (model.py:69-83)."""
    return _embedding("vgg16", "hedngan", pretrained, device, weights_dir)


def gem_resnet101_cyclegan(pretrained=True, device=None, weights_dir=None):
    """GeM ResNet-101, CycleGAN augmentation and CLAHE (model.py:86-100)."""
    return _embedding("resnet101", "cyclegan", pretrained, device, weights_dir)


def gem_resnet101_hedngan(pretrained=True, device=None, weights_dir=None):
    """GeM ResNet-101, HED^N-GAN augmentation and CLAHE (model.py:103-117)."""
    return _embedding("resnet101", "hedngan", pretrained, device, weights_dir)


class _GeneratorNetwork(object):
    """SingleNetwork-shaped holder of a stock-PyTorch generator (no wrappers; generator.yml:15-19)."""

    def __init__(self, model, device):
        from .transforms import initialize_transforms
        self.model = model.to(device).eval()
        self.device = device
        self.meta = model.meta
        self.transform = initialize_transforms("pil2np | totensor | normalize", [[0.5, 0.5, 0.5], [0.5, 0.5, 0.5]], device=device)

    def eval(self):
        self.model.eval()
        return self

    def __call__(self, image):
        return self.model(image.to(self.device))


def _generator(tag, pretrained, device, weights_dir, norm_layer, init_kind):
    device = _device(device)
    model = ResnetGenerator(3, 3, n_blocks=9, norm_layer=norm_layer)
    if pretrained:
        state = torch.load(_weights_file(weights_dir, "%s_generator_X.pth" % tag), map_location="cpu", weights_only=False)
        state = state["net"] if "net" in state else state
        model.load_state_dict(state["model_state"] if "model_state" in state else state)
    else:
        torch.manual_seed(0)                                                      # generator.yml:13
        model.apply(lambda mod: init_weights_p2p(mod, init_kind, 0.2))
    return _GeneratorNetwork(model, device)


def cyclegan(pretrained=True, device=None, weights_dir=None):
    """ResNet CycleGAN day-to-night generator (model.py:125-136); stock PyTorch."""
    return _generator("cyclegan", pretrained, device, weights_dir, "instance", "normal")


def hedngan(pretrained=True, device=None, weights_dir=None):
    """ResNet HED^N-GAN day-to-night generator (model.py:139-154); stock PyTorch."""
    return _generator("hedngan", pretrained, device, weights_dir, "instance" if pretrained else "batch",
                      "normal" if pretrained else "kaiming")

"""torch.hub manifest -- same entry points as the reference's hubconf.py:1-4."""
from gandtr_b200.hub import gem_vgg16_cyclegan, gem_vgg16_hedngan, gem_resnet101_cyclegan, gem_resnet101_hedngan, \
    hedngan, cyclegan  # noqa: F401

dependencies = ["torch", "torchvision"]

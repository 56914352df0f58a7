"""Day-to-night ResNet generators (CycleGAN / HED^N-GAN) in STOCK PyTorch.

The generators are outside the accelerated path (BASELINE.json north_star: "the CycleGAN/HED^N GAN generators stay in
stock PyTorch"); this module only rebuilds the architecture of `official_resnet_generator`
(mdir/components/model/network/p2p_networks.py:239-313,454-506 as configured by mdir/hub/generator.yml:4-10) with the
same `model.<idx>` state_dict layout so reference checkpoints load, and so the in-line GAN -> ClahePost -> embed chain
(BASELINE config 5) can be assembled on the device.
"""
import torch
import torch.nn as nn


def _norm(kind, ch, track_running_stats=True):
    if kind == "instance":
        return nn.InstanceNorm2d(ch, affine=False)
    if kind == "batch":
        return nn.BatchNorm2d(ch, affine=True, track_running_stats=track_running_stats)
    raise NotImplementedError("normalization layer [%s] is not found" % kind)


class ResnetBlock(nn.Module):
    def __init__(self, dim, norm_layer, use_bias):
        super().__init__()
        self.conv_block = nn.Sequential(
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=use_bias), _norm(norm_layer, dim),
            nn.ReLU(True),
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=use_bias), _norm(norm_layer, dim))

    def forward(self, x):
        return x + self.conv_block(x)


class ResnetGenerator(nn.Module):
    def __init__(self, input_nc=3, output_nc=3, ngf=64, norm_layer="instance", n_blocks=9, no_antialias=True,
                 no_antialias_up=True):
        super().__init__()
        if not (no_antialias and no_antialias_up):
            raise NotImplementedError("anti-aliased down/up-sampling is not used by the hub generators (generator.yml:5-6)")
        self.meta = {"in_channels": input_nc, "out_channels": output_nc}
        bias = norm_layer == "instance"
        layers = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=bias),
                  _norm(norm_layer, ngf), nn.ReLU(True)]
        ch = ngf
        for _ in range(2):
            layers += [nn.Conv2d(ch, ch * 2, kernel_size=3, stride=2, padding=1, bias=bias), _norm(norm_layer, ch * 2),
                       nn.ReLU(True)]
            ch *= 2
        layers += [ResnetBlock(ch, norm_layer, bias) for _ in range(n_blocks)]
        for _ in range(2):
            layers += [nn.ConvTranspose2d(ch, ch // 2, kernel_size=3, stride=2, padding=1, output_padding=1, bias=bias),
                       _norm(norm_layer, ch // 2), nn.ReLU(True)]
            ch //= 2
        layers += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0), nn.Tanh()]
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


def init_weights_p2p(module, kind="normal", init_gain=0.2):
    """mdir/components/model/weight_initialization.py:62-76."""
    name = module.__class__.__name__
    if hasattr(module, "weight") and module.weight is not None and ("Conv" in name or "Linear" in name):
        if kind == "normal":
            nn.init.normal_(module.weight.data, 0.0, init_gain)
        else:
            nn.init.kaiming_normal_(module.weight.data, a=0, mode="fan_in")
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias.data, 0.0)
    elif "BatchNorm2d" in name:
        nn.init.normal_(module.weight.data, 1.0, init_gain)
        nn.init.constant_(module.bias.data, 0.0)

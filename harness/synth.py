"""Synthetic Office-Home-shaped inputs and a synthetic ResNet-50-DWT checkpoint.

There is no network: neither the Office-Home images nor ``model_best_gr_4.pth.tar``
(/root/reference/README.md:11) exist, so both are synthesised.
Every tensor is drawn from its own ``torch.Generator`` seeded from (seed, key
name), on CPU, so the same call gives bit-identical tensors in the build
container and on the GPU box regardless of construction order.
"""
from __future__ import annotations

import zlib

import torch

_STAGES = ((64, 3), (128, 4), (256, 6), (512, 3))
EXPANSION = 4


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1_000_003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def _randn(seed, key, *shape):
    return torch.randn(*shape, generator=_gen(seed, key), dtype=torch.float32)


def _rand(seed, key, *shape):
    return torch.rand(*shape, generator=_gen(seed, key), dtype=torch.float32)


def _spd(seed, key, groups, gs):
    a = _randn(seed, key, groups, gs, gs)
    return a @ a.transpose(1, 2) / gs + 0.5 * torch.eye(gs)


def norm_site_keys():
    """(key prefix, channels, whitening?) for every norm site of the model, in the
    checkpoint's naming (resnet50_dwt_mec_officehome.py:76-213,274-288)."""
    sites = [("bn1", 64, True)]
    for li, (planes, blocks) in enumerate(_STAGES, start=1):
        for b in range(blocks):
            p = f"layer{li}.{b}"
            sites += [(p + ".bn1", planes, li == 1), (p + ".bn2", planes, li == 1),
                      (p + ".bn3", planes * EXPANSION, li == 1)]
        sites.append((f"layer{li}.0.downsample_bn", planes * EXPANSION, li == 1))
    return sites


def synth_state_dict(seed: int = 1, group_size: int = 4, with_convs: bool = True,
                     num_classes: int = 65) -> dict:
    """Checkpoint contents with the 'module.' prefix already stripped."""
    sd = {}
    for key, c, whiten in norm_site_keys():
        if whiten:
            sd[key + ".wh.running_mean"] = 0.1 * _randn(seed, key + "rm", 1, c, 1, 1)
            sd[key + ".wh.running_variance"] = _spd(seed, key + "rv", c // group_size, group_size)
            sd[key + ".gamma"] = 0.5 + _rand(seed, key + "g", c, 1, 1)
            sd[key + ".beta"] = 0.1 * _randn(seed, key + "b", c, 1, 1)
        else:
            sd[key + ".running_mean"] = 0.1 * _randn(seed, key + "rm", c)
            sd[key + ".running_var"] = 0.5 + _rand(seed, key + "rv", c)
            sd[key + ".weight"] = 0.5 + _rand(seed, key + "g", c)
            sd[key + ".bias"] = 0.1 * _randn(seed, key + "b", c)
    if with_convs:
        def conv(name, cout, cin, k):
            std = (2.0 / (cout * k * k)) ** 0.5                 # kaiming_normal_, fan_out, relu
            sd[name + ".weight"] = std * _randn(seed, name, cout, cin, k, k)
        conv("conv1", 64, 3, 7)
        inplanes = 64
        for li, (planes, blocks) in enumerate(_STAGES, start=1):
            for b in range(blocks):
                p = f"layer{li}.{b}"
                conv(p + ".conv1", planes, inplanes, 1)
                conv(p + ".conv2", planes, planes, 3)
                conv(p + ".conv3", planes * EXPANSION, planes, 1)
                if b == 0:
                    conv(p + ".downsample.0", planes * EXPANSION, inplanes, 1)
                inplanes = planes * EXPANSION
        bound = (1.0 / (512 * EXPANSION)) ** 0.5
        sd["fc_out.weight"] = (2 * _rand(seed, "fcw", num_classes, 512 * EXPANSION) - 1) * bound
        sd["fc_out.bias"] = (2 * _rand(seed, "fcb", num_classes) - 1) * bound
    return sd


def synth_batch(seed: int, per_domain: int, size: int = 224, num_classes: int = 65):
    """(images [3*per_domain,3,size,size], source labels [per_domain]): source | target | target-aug,
    the concatenation order of resnet50_dwt_mec_officehome.py:416."""
    x = _randn(seed, "images", 3 * per_domain, 3, size, size)
    y = torch.randint(0, num_classes, (per_domain,), generator=_gen(seed, "labels"))
    return x, y

"""The reference UNet's 136-entry ``state_dict`` contract (code/model.py:53-68)
and helpers around it.  Pure host logic -- no kernels here.

Key layout, for every DoubleConvLayer prefix P in
``downconv{1-4}.conv``, ``bottleneck``, ``upconv{1-4}.conv``::

    P.double_conv.{0,3}.weight  (Co, Ci, 3, 3)   P.double_conv.{0,3}.bias (Co,)
    P.double_conv.{1,4}.{weight,bias,running_mean,running_var} (Co,)
    P.double_conv.{1,4}.num_batches_tracked ()  int64
    upconv{1-4}.up.weight (Ci, Co, 2, 2)   upconv{1-4}.up.bias (Co,)
    out.weight (1, 64, 1, 1)   out.bias (1,)
"""
from __future__ import annotations

from collections import OrderedDict

import torch

BN_EPS = 1e-5        # nn.BatchNorm2d default, code/model.py:12,15
BN_MOMENTUM = 0.1

# (attribute prefix, in_channels, out_channels) in module-definition order (model.py:56-68)
DOWN_BLOCKS = [("downconv1", 1, 64), ("downconv2", 64, 128), ("downconv3", 128, 256), ("downconv4", 256, 512)]
BOTTLENECK = ("bottleneck", 512, 1024)
UP_BLOCKS = [("upconv1", 1024, 512), ("upconv2", 512, 256), ("upconv3", 256, 128), ("upconv4", 128, 64)]


def double_conv_prefixes():
    """[(prefix, Ci, Co)] of every DoubleConvLayer in state_dict order."""
    out = [(f"{n}.conv", ci, co) for n, ci, co in DOWN_BLOCKS]
    out.append(BOTTLENECK)
    out += [(f"{n}.conv", ci, co) for n, ci, co in UP_BLOCKS]
    return out


def state_dict_spec(in_channels: int = 1, num_classes: int = 1) -> "OrderedDict[str, tuple]":
    """Ordered {key: (shape, dtype)} exactly as ``UNet().state_dict()`` yields it."""
    spec: "OrderedDict[str, tuple]" = OrderedDict()

    def dconv(p, ci, co):
        for conv_i, bn_i, cin in ((0, 1, ci), (3, 4, co)):
            spec[f"{p}.double_conv.{conv_i}.weight"] = ((co, cin, 3, 3), torch.float32)
            spec[f"{p}.double_conv.{conv_i}.bias"] = ((co,), torch.float32)
            for leaf in ("weight", "bias", "running_mean", "running_var"):
                spec[f"{p}.double_conv.{bn_i}.{leaf}"] = ((co,), torch.float32)
            spec[f"{p}.double_conv.{bn_i}.num_batches_tracked"] = ((), torch.int64)

    for n, ci, co in DOWN_BLOCKS:
        dconv(f"{n}.conv", in_channels if n == "downconv1" else ci, co)
    dconv(*BOTTLENECK)
    for n, ci, co in UP_BLOCKS:
        spec[f"{n}.up.weight"] = ((ci, co, 2, 2), torch.float32)
        spec[f"{n}.up.bias"] = ((co,), torch.float32)
        dconv(f"{n}.conv", ci, co)
    spec["out.weight"] = ((num_classes, 64, 1, 1), torch.float32)
    spec["out.bias"] = ((num_classes,), torch.float32)
    return spec


def seeded_state_dict(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic random-init checkpoint in the reference layout (the reference ships
    no checkpoint; test.py:59 expects ``./saved_models/unet_denoiser_{noise}.pth``).

    He-style conv weights and *randomised* BatchNorm affine / running statistics, so that
    eval-mode BN folding is actually exercised (default-init BN is the identity).  Drawn
    from a CPU ``torch.Generator`` so the same seed gives the same tensors on every box.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, (shape, dtype) in state_dict_spec().items():
        leaf = key.rsplit(".", 1)[1]
        if dtype == torch.int64:
            sd[key] = torch.tensor(100, dtype=torch.int64)
        elif len(shape) == 4:                       # conv / conv-transpose weight
            if ".up." in key:
                fan_in = shape[0]                   # each output pixel sees Ci inputs through one tap
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            std = (2.0 / fan_in) ** 0.5
            sd[key] = torch.randn(shape, generator=g, dtype=torch.float32) * std
        elif leaf == "running_var":
            sd[key] = 0.5 + torch.rand(shape, generator=g, dtype=torch.float32)
        elif leaf == "running_mean":
            sd[key] = 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        elif leaf == "weight":                      # BN gamma
            sd[key] = 0.75 + 0.5 * torch.rand(shape, generator=g, dtype=torch.float32)
        else:                                       # conv bias / BN beta
            sd[key] = 0.05 * torch.randn(shape, generator=g, dtype=torch.float32)
    return sd


def check_state_dict(sd) -> None:
    """Raise the same kind of error ``load_state_dict(strict=True)`` would for a dict that
    does not match the reference layout."""
    spec = state_dict_spec()
    missing = [k for k in spec if k not in sd]
    unexpected = [k for k in sd if k not in spec]
    if missing or unexpected:
        raise RuntimeError(f"Error(s) in loading state_dict for UNet: missing keys {missing}, unexpected keys {unexpected}")
    for k, (shape, _dtype) in spec.items():
        if tuple(sd[k].shape) != tuple(shape):
            raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(sd[k].shape)} vs model {tuple(shape)}")

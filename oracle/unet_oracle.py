"""fp32 CPU restatement of the reference UNet forward (code/model.py:7-94), eval-mode
BatchNorm, as a pure function of a reference-layout ``state_dict``.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Pinned: tests/test_oracle_unet.py checks it against tests/golden/unet_*.npz, which
oracle/make_golden.py produced by running the reference's own ``model.UNet`` imported
from /root/reference/code (that tree does not exist on the GPU box, hence the fixtures).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _double_conv(x, sd, p):
    """DoubleConvLayer.forward, model.py:10-20: (conv3x3 pad1 -> BN(eval) -> ReLU) x 2."""
    for conv_i, bn_i in ((0, 1), (3, 4)):
        x = F.conv2d(x, sd[f"{p}.double_conv.{conv_i}.weight"], sd[f"{p}.double_conv.{conv_i}.bias"], padding=1)
        x = F.batch_norm(x, sd[f"{p}.double_conv.{bn_i}.running_mean"], sd[f"{p}.double_conv.{bn_i}.running_var"],
                         sd[f"{p}.double_conv.{bn_i}.weight"], sd[f"{p}.double_conv.{bn_i}.bias"],
                         training=False, eps=BN_EPS)
        x = F.relu(x)
    return x


def _up(x1, x2, sd, name):
    """UpSampleLayer.forward, model.py:41-50: ConvT k2 s2 -> pad to skip size -> cat[skip, up] -> DoubleConv."""
    x1 = F.conv_transpose2d(x1, sd[f"{name}.up.weight"], sd[f"{name}.up.bias"], stride=2)
    dy = x2.shape[2] - x1.shape[2]
    dx = x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return _double_conv(torch.cat([x2, x1], dim=1), sd, f"{name}.conv")


@torch.no_grad()
def unet_forward(sd, x: torch.Tensor, return_intermediates: bool = False):
    """UNet.forward, model.py:70-94.  x: (N,1,F,T) float32 on CPU."""
    sd = {k: v.to(torch.float32) if v.is_floating_point() else v for k, v in sd.items()}
    x = x.to(torch.float32)
    inter = {}
    skips = []
    h = x
    for i in range(1, 5):
        s = _double_conv(h, sd, f"downconv{i}.conv")          # DownSampleLayer.forward model.py:29-32
        skips.append(s)
        inter[f"down{i}"] = s
        h = F.max_pool2d(s, 2)
    h = _double_conv(h, sd, "bottleneck")
    inter["bottle"] = h
    for i in range(1, 5):
        h = _up(h, skips[4 - i], sd, f"upconv{i}")
        inter[f"up{i}"] = h
    out = F.conv2d(h, sd["out.weight"], sd["out.bias"])       # model.py:68,93 -- no activation
    if return_intermediates:
        return out, inter
    return out

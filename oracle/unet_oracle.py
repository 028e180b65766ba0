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


def _r(t, dtype=torch.bfloat16):
    """Round to bfloat16 and back: the storage / operand precision of the B200 build's activations and GEMM weights.  (float16 is
    accepted for what-if studies of the recipe: DESIGN.md, Numerics.)"""
    return t.to(dtype).to(torch.float32)


def _double_conv(x, sd, p, bf16=False, first_fp32=False, keep_last_fp32=False, wdt=torch.bfloat16):
    """DoubleConvLayer.forward, model.py:10-20: (conv3x3 pad1 -> BN(eval) -> ReLU) x 2.
    bf16=True emulates the B200 recipe in fp32 arithmetic: conv weights rounded to bf16 (except the Cin=1 first layer, which
    the build computes in fp32), every activation rounded to bf16 where the build stores it, BN / ReLU in fp32."""
    for conv_i, bn_i in ((0, 1), (3, 4)):
        w = sd[f"{p}.double_conv.{conv_i}.weight"]
        if bf16 and not (first_fp32 and conv_i == 0):
            w = _r(w, wdt)
        x = F.conv2d(x, w, sd[f"{p}.double_conv.{conv_i}.bias"], padding=1)
        x = F.batch_norm(x, sd[f"{p}.double_conv.{bn_i}.running_mean"], sd[f"{p}.double_conv.{bn_i}.running_var"],
                         sd[f"{p}.double_conv.{bn_i}.weight"], sd[f"{p}.double_conv.{bn_i}.bias"],
                         training=False, eps=BN_EPS)
        x = F.relu(x)
        if bf16 and not (keep_last_fp32 and conv_i == 3):
            x = _r(x)
    return x


def _up(x1, x2, sd, name, bf16=False, keep_last_fp32=False, wdt=torch.bfloat16):
    """UpSampleLayer.forward, model.py:41-50: ConvT k2 s2 -> pad to skip size -> cat[skip, up] -> DoubleConv."""
    wt = sd[f"{name}.up.weight"]
    x1 = F.conv_transpose2d(x1, _r(wt, wdt) if bf16 else wt, sd[f"{name}.up.bias"], stride=2)
    if bf16:
        x1 = _r(x1)
    dy = x2.shape[2] - x1.shape[2]
    dx = x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return _double_conv(torch.cat([x2, x1], dim=1), sd, f"{name}.conv", bf16=bf16, keep_last_fp32=keep_last_fp32, wdt=wdt)


@torch.no_grad()
def unet_forward(sd, x: torch.Tensor, return_intermediates: bool = False, emulate_bf16: bool = False, weight_format: str = "bf16"):
    """UNet.forward, model.py:70-94.  x: (N,1,F,T) float32 on CPU.

    emulate_bf16=True is NOT the reference: it is the reference's graph with the B200 build's roundings inserted (bf16 GEMM
    operands and stored activations, fp32 accumulation / BatchNorm / ReLU / head), used by the tests to separate "the kernels
    implement the stated recipe" (GPU vs this emulation: tight) from "what the recipe costs" (emulation vs fp32: the 1e-2 budget)."""
    sd = {k: v.to(torch.float32) if v.is_floating_point() else v for k, v in sd.items()}
    x = x.to(torch.float32)
    bf = bool(emulate_bf16)
    wdt = torch.float16 if weight_format == "f16" else torch.bfloat16
    inter = {}
    skips = []
    h = x
    for i in range(1, 5):
        s = _double_conv(h, sd, f"downconv{i}.conv", bf16=bf, first_fp32=(i == 1), wdt=wdt)          # DownSampleLayer.forward model.py:29-32
        skips.append(s)
        inter[f"down{i}"] = s
        h = F.max_pool2d(s, 2)
    h = _double_conv(h, sd, "bottleneck", bf16=bf, wdt=wdt)
    inter["bottle"] = h
    for i in range(1, 5):
        h = _up(h, skips[4 - i], sd, f"upconv{i}", bf16=bf, keep_last_fp32=(i == 4), wdt=wdt)        # the fused 1x1 head reads fp32 accumulators
        inter[f"up{i}"] = h
    out = F.conv2d(h, sd["out.weight"], sd["out.bias"])       # model.py:68,93 -- no activation
    if return_intermediates:
        return out, inter
    return out

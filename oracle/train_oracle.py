"""fp32 CPU restatement of ONE iteration of the reference's train_one_epoch body (code/train.py:65-72):
zero_grad -> model(noisy) in train() mode (code/model.py:70-94 with batch-statistics BatchNorm) -> CombinedPerceptualLoss
(code/loss.py:83-95) -> backward -> clip_grad_norm_(max_norm=1.0) -> AdamW(lr) step.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Pinned: tests/test_oracle_train.py checks it against tests/golden/train_step_small.npz, produced by
oracle/make_golden_train.py from the reference's own ``model.UNet`` / ``loss.CombinedPerceptualLoss`` + torch.optim.AdamW.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import loss_oracle

BN_EPS, BN_MOMENTUM = 1e-5, 0.1
_BUFFER_LEAVES = ("running_mean", "running_var", "num_batches_tracked")


def _double_conv_train(x, p, b, prefix):
    for conv_i, bn_i in ((0, 1), (3, 4)):
        x = F.conv2d(x, p[f"{prefix}.double_conv.{conv_i}.weight"], p[f"{prefix}.double_conv.{conv_i}.bias"], padding=1)
        bn = f"{prefix}.double_conv.{bn_i}"
        x = F.batch_norm(x, b[f"{bn}.running_mean"], b[f"{bn}.running_var"], p[f"{bn}.weight"], p[f"{bn}.bias"], training=True,
                         momentum=BN_MOMENTUM, eps=BN_EPS)          # updates the running estimates in place, like nn.BatchNorm2d
        b[f"{bn}.num_batches_tracked"] += 1
        x = F.relu(x)
    return x


def unet_forward_train(p, b, x):
    """UNet.forward (model.py:70-94) in train() mode.  p: parameters, b: buffers (updated in place)."""
    skips = []
    h = x
    for i in range(1, 5):
        s = _double_conv_train(h, p, b, f"downconv{i}.conv")
        skips.append(s)
        h = F.max_pool2d(s, 2)
    h = _double_conv_train(h, p, b, "bottleneck")
    for i in range(1, 5):
        up = F.conv_transpose2d(h, p[f"upconv{i}.up.weight"], p[f"upconv{i}.up.bias"], stride=2)
        sk = skips[4 - i]
        dy, dx = sk.shape[2] - up.shape[2], sk.shape[3] - up.shape[3]
        up = F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        h = _double_conv_train(torch.cat([sk, up], dim=1), p, b, f"upconv{i}.conv")
    return F.conv2d(h, p["out.weight"], p["out.bias"])


def split_state_dict(sd):
    params, buffers = OrderedDict(), OrderedDict()
    for k, v in sd.items():
        if k.rsplit(".", 1)[1] in _BUFFER_LEAVES:
            buffers[k] = v.clone()
        else:
            params[k] = v.detach().clone().float().requires_grad_()
    return params, buffers


class TrainOracle:
    """Holds parameters, buffers and a torch AdamW (train.py:124: AdamW(model.parameters(), lr) with torch defaults)."""

    def __init__(self, sd, lr=1e-4, max_norm=1.0):
        self.params, self.buffers = split_state_dict(sd)
        self.opt = torch.optim.AdamW(list(self.params.values()), lr=lr)
        self.max_norm = max_norm

    def forward_backward(self, noisy, clean):
        self.opt.zero_grad()
        out = unet_forward_train(self.params, self.buffers, noisy.float())
        total, stft, mel, l1 = loss_oracle.combined_loss(out, clean.float())
        total.backward()
        return out.detach(), torch.stack([total.detach(), stft.detach(), mel.detach(), l1.detach()])

    def step(self):
        norm = torch.nn.utils.clip_grad_norm_(list(self.params.values()), max_norm=self.max_norm)
        self.opt.step()
        return norm

    def train_step(self, noisy, clean):
        out, losses = self.forward_backward(noisy, clean)
        grads = OrderedDict((k, v.grad.detach().clone()) for k, v in self.params.items())
        norm = self.step()
        return out, losses, grads, norm

    def state_dict(self):
        sd = OrderedDict((k, v.detach().clone()) for k, v in self.params.items())
        sd.update((k, v.clone()) for k, v in self.buffers.items())
        return sd

#!/usr/bin/env python
"""Per-layer error of the train-mode forward against a CPU fp32 forward (debug)."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodenoiser_b200.checkpoint import seeded_state_dict
from audiodenoiser_b200.model import UNet
from audiodenoiser_b200.training import TrainEngine
from oracle.make_golden_train import batch
from oracle.train_oracle import split_state_dict

shape = (4, 1, 256, 64)
net = UNet(); net.load_state_dict(seeded_state_dict(7))
eng = TrainEngine(net, device=torch.device("cuda", 0))
noisy, clean = batch(101, shape)
out = eng.forward(noisy.cuda())
sv = eng.saved
p, b = split_state_dict(seeded_state_dict(7))
ref = {}
def dc(x, prefix):
    for ci, bi in ((0, 1), (3, 4)):
        z = F.conv2d(x, p[f"{prefix}.double_conv.{ci}.weight"], p[f"{prefix}.double_conv.{ci}.bias"], padding=1)
        bn = f"{prefix}.double_conv.{bi}"
        x = F.relu(F.batch_norm(z, None, None, p[f"{bn}.weight"], p[f"{bn}.bias"], training=True, eps=1e-5))
        ref[(prefix, ci)] = (z.detach(), x.detach())
    return x
with torch.no_grad():
    skips = []; h = noisy
    for i in range(1, 5):
        s = dc(h, f"downconv{i}.conv"); skips.append(s); h = F.max_pool2d(s, 2)
    h = dc(h, "bottleneck")
    for i in range(1, 5):
        up = F.conv_transpose2d(h, p[f"upconv{i}.up.weight"], p[f"upconv{i}.up.bias"], stride=2)
        ref[f"upconv{i}.up"] = up
        h = dc(torch.cat([skips[4 - i], up], 1), f"upconv{i}.conv")
    o = F.conv2d(h, p["out.weight"], p["out.bias"])
nchw = lambda t: t.float().cpu().permute(0, 3, 1, 2)
e = lambda a, r: float((a - r).norm() / r.norm())
for layer in eng.layers:
    key = (layer[0], layer[1])
    z, y, _, _ = sv[key]
    zr, yr = ref[key]
    print(f"{key[0]}.{key[1]}: z err {e(nchw(z), zr):.4f}  y err {e(nchw(y), yr):.4f}  |mean/std| of z: {float((zr.mean((0,2,3)).abs() / zr.std((0,2,3))).mean()):.2f}")
for i in range(1, 5):
    print(f"upconv{i}.up err {e(nchw(sv[f'upconv{i}.up'][1]), ref[f'upconv{i}.up']):.4f}")
print("out err", e(out.cpu(), o))

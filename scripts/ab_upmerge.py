#!/usr/bin/env python
"""In-process A/B of the merged ConvTranspose+conv kernel (csrc/conv_upm.cu) against the two-kernel path: per-layer times of the
UNet forward at the bench shape, interleaved.  usage: python scripts/ab_upmerge.py [batch]"""
import sys

import torch

sys.path.insert(0, ".")
from audiodenoiser_b200.checkpoint import seeded_state_dict
from audiodenoiser_b200.model import UNet

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = UNet().eval()
net.load_state_dict(seeded_state_dict(3))
x = torch.rand(batch, 1, 257, 1034, device="cuda")
MODES = {"two-kernel": (False, False), "merged": (True, False), "merged+plane128": (True, True)}
res = {m: {} for m in MODES}
outs = {}
with torch.no_grad():
    for rep in range(6):
        for mode, flag in MODES.items():
            net.upmerge, net.plane128 = flag
            net.profile = []
            y = net(x)
            torch.cuda.synchronize()
            outs[mode] = y[:2].float().clone()
            if rep >= 2:
                for layer, kind, flops, a, b in net.profile:
                    d = res[mode].setdefault(layer, [0.0, 0, flops])
                    d[0] += a.elapsed_time(b); d[1] += 1
            net.profile = None
tot = {}
for m in MODES:
    tot[m] = sum(v[0] / v[1] for v in res[m].values())
    print(m, "  ".join(f"{k} {v[0] / v[1]:.3f} ({v[2] / (v[0] / v[1]) / 1e9:.0f} TF)" for k, v in res[m].items() if k.startswith("upconv") or k.startswith("downconv2")))
print("forward total: " + ", ".join(f"{m} {tot[m]:.3f} ms" for m in MODES))
a, b, c = outs["two-kernel"], outs["merged"], outs["merged+plane128"]
print(f"merged vs two-kernel output: norm-rel {((a - b).norm() / a.norm()).item():.3e}; plane128 vs merged: max abs {(b - c).abs().max().item():.3e}")

#!/usr/bin/env python
"""In-process A/B of conv_halo variants (debug hooks adn__conv_halo_pitch / adn__conv_pair_mode): per-layer times of the UNet
forward at the bench shape, interleaved so box-to-box clock variance cancels.  usage: python scripts/ab_halo.py [batch]"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from audiodenoiser_b200 import _lib
from audiodenoiser_b200.checkpoint import seeded_state_dict
from audiodenoiser_b200.model import UNet

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = _lib.load()
for f in (lib.adn__conv_halo_pitch, lib.adn__conv_pair_mode, lib.adn__conv_halo_stages, lib.adn__conv_halo_tune):
    f.argtypes = [ctypes.c_int]; f.restype = None
net = UNet().eval()
net.load_state_dict(seeded_state_dict(3))
x = torch.rand(batch, 1, 257, 1034, device="cuda")
MODES = {"default": (10, 3, 0, 2), "per-warp-stores": (10, 3, 0, 6)}
res = {m: {} for m in MODES}
outs = {}
with torch.no_grad():
    for rep in range(6):
        for mode, (pitch, pair, stages, tune) in MODES.items():
            lib.adn__conv_halo_pitch(pitch); lib.adn__conv_pair_mode(pair); lib.adn__conv_halo_stages(stages); lib.adn__conv_halo_tune(tune)
            net.profile = []
            y = net(x)
            torch.cuda.synchronize()
            outs[mode] = y[:2].float().clone()
            if rep >= 2:
                for layer, kind, flops, a, b in net.profile:
                    d = res[mode].setdefault(layer, [0.0, 0, flops])
                    d[0] += a.elapsed_time(b); d[1] += 1
            net.profile = None
base = next(iter(MODES))
tot = {m: 0.0 for m in MODES}
for layer in res[base]:
    ms = {m: res[m][layer][0] / res[m][layer][1] for m in MODES}
    for m in MODES:
        tot[m] += ms[m]
    fl = res[base][layer][2]
    if max(ms.values()) / min(ms.values()) > 1.015:
        print(f"{layer:16s} " + "   ".join(f"{m} {ms[m]:7.3f} ms ({fl / ms[m] / 1e9:5.0f} TF)" for m in MODES))
print("forward total: " + ", ".join(f"{m} {tot[m]:.3f} ms" for m in MODES))
for m in MODES:
    print(f"max |out - out[{base}]| for {m}: {(outs[m] - outs[base]).abs().max().item():.3e}")

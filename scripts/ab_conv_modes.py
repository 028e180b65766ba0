#!/usr/bin/env python
"""In-process A/B of the 64-output-channel conv kernels (debug hook adn__conv_dx_mode): per-layer times of the UNet forward at
the bench shape with the kx-in-N kernel (conv_dx.cu) on and off, interleaved so box-to-box clock variance cancels.
usage: python scripts/ab_conv_modes.py [batch]"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from audiodenoiser_b200 import _lib
from audiodenoiser_b200.checkpoint import seeded_state_dict
from audiodenoiser_b200.model import UNet

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = _lib.load()
lib.adn__conv_dx_mode.argtypes = [ctypes.c_int]
lib.adn__conv_dx_mode.restype = None
net = UNet().eval()
net.load_state_dict(seeded_state_dict(3))
x = torch.rand(batch, 1, 257, 1034, device="cuda")
MODES = (0, 4, 1)
res = {m: {} for m in MODES}
with torch.no_grad():
    for rep in range(6):
        for mode in MODES:
            lib.adn__conv_dx_mode(mode)
            net.profile = []
            net(x)
            torch.cuda.synchronize()
            if rep >= 2:
                for layer, kind, flops, a, b in net.profile:
                    d = res[mode].setdefault(layer, [0.0, 0, flops])
                    d[0] += a.elapsed_time(b); d[1] += 1
            net.profile = None
tot = {m: 0.0 for m in MODES}
names = {0: "halo", 1: "dx-auto", 2: "dx-pair", 3: "dx-4sets", 4: "dx-2sets"}
for layer in res[0]:
    ms = {m: res[m][layer][0] / res[m][layer][1] for m in MODES}
    for m in MODES:
        tot[m] += ms[m]
    fl = res[0][layer][2]
    if max(ms.values()) / min(ms.values()) > 1.03:
        print(f"{layer:16s} " + "   ".join(f"{names[m]} {ms[m]:7.3f} ms ({fl / ms[m] / 1e9:5.0f} TF)" for m in MODES))
print("forward total: " + ", ".join(f"{names[m]} {tot[m]:.3f} ms" for m in MODES))

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
res = {0: {}, 1: {}}
with torch.no_grad():
    for rep in range(6):
        for mode in (0, 1):
            lib.adn__conv_dx_mode(mode)
            net.profile = []
            net(x)
            torch.cuda.synchronize()
            if rep >= 2:
                for layer, kind, flops, a, b in net.profile:
                    d = res[mode].setdefault(layer, [0.0, 0, flops])
                    d[0] += a.elapsed_time(b); d[1] += 1
            net.profile = None
tot = {0: 0.0, 1: 0.0}
for layer in res[0]:
    m0 = res[0][layer][0] / res[0][layer][1]; m1 = res[1][layer][0] / res[1][layer][1]
    tot[0] += m0; tot[1] += m1
    fl = res[0][layer][2]
    if abs(m0 - m1) / m0 > 0.03:
        print(f"{layer:16s} halo {m0:7.3f} ms ({fl / m0 / 1e9:6.0f} TF)   dx {m1:7.3f} ms ({fl / m1 / 1e9:6.0f} TF)")
print(f"forward total: halo {tot[0]:.3f} ms, dx {tot[1]:.3f} ms")

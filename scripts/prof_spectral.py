#!/usr/bin/env python
"""Driver for ncu captures of the spectral kernels at BASELINE config 2 (4 096 clips, 3 s @ 8 kHz): 3 STFT launches, 3 iSTFT
launches with the in-kernel phase and 3 with an explicit phasor.
usage: ncu --set full -k regex:stft_ws_kernel -s 2 -c 1 ... python scripts/prof_spectral.py [n_clips]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodenoiser_b200 import spectral  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
x = torch.rand((n, 24000), device=dev) * 2 - 1
mag = torch.empty((n, 257, 188), device=dev)
audio = torch.empty((n, 128 * 187), device=dev)
for i in range(3):
    spectral.stft_mag_batched(x, True, out=mag)
for i in range(3):
    spectral.istft_batched(mag, None, seed=i, out=audio)
ph = torch.polar(torch.ones_like(mag), torch.rand_like(mag) * 6.2831853)
for i in range(3):
    spectral.istft_batched(mag, ph, out=audio)
torch.cuda.synchronize()
print("ok")

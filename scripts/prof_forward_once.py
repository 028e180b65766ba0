#!/usr/bin/env python
"""One eval forward of the UNet at the bench shape (for ncu captures of single layers).  usage: python scripts/prof_forward_once.py [batch]"""
import sys

import torch

sys.path.insert(0, ".")
from audiodenoiser_b200.checkpoint import seeded_state_dict
from audiodenoiser_b200.model import UNet

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = UNet().eval()
net.load_state_dict(seeded_state_dict(3))
x = torch.rand(batch, 1, 257, 1034, device="cuda")
with torch.no_grad():
    for _ in range(2):
        net(x)
torch.cuda.synchronize()
print("ok")

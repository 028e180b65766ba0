#!/usr/bin/env python
"""Driver for ncu launch lists of the training step: 1 warm + 2 eager steps at batch 16 x (1,256,64)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodenoiser_b200.checkpoint import seeded_state_dict
from audiodenoiser_b200.model import UNet
from audiodenoiser_b200.training import TrainEngine
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net = UNet(); net.load_state_dict(seeded_state_dict(7))
eng = TrainEngine(net, device=torch.device("cuda", 0))
g = torch.Generator().manual_seed(1)
clean = (torch.rand((b, 1, 256, 64), generator=g) * 2).cuda(); noisy = (clean + 0.3 * torch.rand((b, 1, 256, 64), generator=g).cuda())
for i in range(3):
    l = eng.train_step(noisy, clean)
torch.cuda.synchronize()
print("ok", l.tolist())

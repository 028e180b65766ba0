"""Generate tests/golden/train_step_small.npz from the UNMODIFIED reference: code/model.py UNet in train() mode +
code/loss.py CombinedPerceptualLoss + the train_one_epoch body of code/train.py:65-72 (torch.optim.AdamW(lr=1e-4),
clip_grad_norm_(1.0)), two consecutive steps on a seeded (2,1,32,32) batch.  train.py itself cannot be imported (it imports
modules that do not exist in the tree), so its six-line body is executed literally here.  Run in the build container
only (/root/reference is absent on the GPU box).  TEST INFRASTRUCTURE.

A 31 M-parameter gradient does not fit a fixture: per tensor we keep the L2 norm, the sum and 8 evenly spaced entries of the
gradient (step 1) and of the parameter after each step."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = "/root/reference/code"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from audiodenoiser_b200.checkpoint import seeded_state_dict  # noqa: E402


def sample_idx(numel):
    return np.unique(np.linspace(0, numel - 1, 8).astype(np.int64))


def batch(seed, shape=(2, 1, 32, 32)):
    g = torch.Generator().manual_seed(seed)
    clean = (torch.randn(shape, generator=g).abs() * torch.rand(shape, generator=g) * 2.0).float()
    noisy = (clean + 0.3 * torch.randn(shape, generator=g).abs()).float()
    return noisy.half().float(), clean.half().float()        # the loader's float16 round trip (data_loader.py:41-42)


def main():
    import model as ref_model
    import loss as ref_loss
    torch.manual_seed(0)
    net = ref_model.UNet(in_channels=1, num_classes=1)
    net.load_state_dict(seeded_state_dict(7), strict=True)
    net.train()
    criterion = ref_loss.CombinedPerceptualLoss()
    optimizer = torch.optim.AdamW(net.parameters(), lr=1e-4)       # train.py:124
    out = {}
    for step in (1, 2):
        noisy, clean = batch(100 + step)
        optimizer.zero_grad()                                        # train.py:66
        outputs = net(noisy)                                         # :67
        loss, ls, lm, l1 = criterion(outputs, clean)                 # :68
        loss.backward()                                              # :69
        if step == 1:
            for k, p in net.named_parameters():
                g = p.grad.detach().reshape(-1)
                out[f"grad_norm/{k}"] = np.float64(g.double().norm().item())
                out[f"grad_sum/{k}"] = np.float64(g.double().sum().item())
                out[f"grad_samples/{k}"] = g[sample_idx(g.numel())].numpy()
        norm = torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)   # :70
        optimizer.step()                                             # :71
        out[f"losses/{step}"] = np.array([loss.item(), ls.item(), lm.item(), l1.item()], np.float64)
        out[f"total_norm/{step}"] = np.float64(norm.item())
        out[f"outputs/{step}"] = outputs.detach().numpy()
        for k, v in net.state_dict().items():
            flat = v.detach().reshape(-1)
            out[f"state_samples/{step}/{k}"] = flat[sample_idx(flat.numel())].numpy()
    path = os.path.join(ROOT, "tests", "golden", "train_step_small.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; losses", out["losses/1"], out["losses/2"], "norms", out["total_norm/1"], out["total_norm/2"])


if __name__ == "__main__":
    main()

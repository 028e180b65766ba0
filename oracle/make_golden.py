"""Generate tests/golden/*.npz from the UNMODIFIED reference (code/model.py, code/loss.py imported
from /root/reference/code).  Run in the build container only -- /root/reference does not exist on
the GPU box, which is why the outputs are committed.  TEST INFRASTRUCTURE.

    python oracle/make_golden.py
"""
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


def magnitude_like(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g).abs() * torch.rand(shape, generator=g) * 3.0).float()


def main():
    import model as ref_model      # /root/reference/code/model.py
    import loss as ref_loss        # /root/reference/code/loss.py

    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- UNet forward (model.py:70-94), eval mode, seeded reference-layout checkpoint ----
    net = ref_model.UNet(in_channels=1, num_classes=1)
    sd = seeded_state_dict(7)
    net.load_state_dict(sd, strict=True)
    net.eval()
    assert list(net.state_dict().keys()) == list(sd.keys())
    for name, shape, seed in (("small", (2, 1, 37, 26), 11), ("train", (2, 1, 256, 64), 12), ("test", (1, 1, 257, 188), 13)):
        x = magnitude_like(shape, seed)
        with torch.no_grad():
            y = net(x)
        np.savez_compressed(os.path.join(gold, f"unet_{name}.npz"), x=x.numpy(), y=y.numpy(), seed=np.int64(7))
        print("unet", name, tuple(y.shape), float(y.abs().max()))

    # ---- CombinedPerceptualLoss (loss.py:83-95) ----
    crit = ref_loss.CombinedPerceptualLoss()
    for name, shape, seed in (("train", (4, 1, 256, 64), 21), ("test", (3, 1, 257, 188), 22)):
        p = magnitude_like(shape, seed)
        t = magnitude_like(shape, seed + 100)
        with torch.no_grad():
            vals = [float(v) for v in crit(p, t)]
        np.savez_compressed(os.path.join(gold, f"loss_{name}.npz"), pred=p.numpy().astype(np.float16),
                            target=t.numpy().astype(np.float16), values=np.array(vals, dtype=np.float64))
        # inputs are stored as float16 to keep the fixture small; recompute on the rounded inputs
        p16 = torch.from_numpy(p.numpy().astype(np.float16)).float()
        t16 = torch.from_numpy(t.numpy().astype(np.float16)).float()
        with torch.no_grad():
            vals = [float(v) for v in crit(p16, t16)]
        np.savez_compressed(os.path.join(gold, f"loss_{name}.npz"), pred=p16.numpy().astype(np.float16),
                            target=t16.numpy().astype(np.float16), values=np.array(vals, dtype=np.float64))
        print("loss", name, vals)

    # SURVEY Appendix A regression value: torch.manual_seed(0); rand(4,1,256,64) x2
    torch.manual_seed(0)
    p, t = torch.rand(4, 1, 256, 64), torch.rand(4, 1, 256, 64)
    with torch.no_grad():
        vals = [float(v) for v in crit(p, t)]
    np.savez_compressed(os.path.join(gold, "loss_seed0.npz"), values=np.array(vals, dtype=np.float64))
    print("loss seed0", vals)


if __name__ == "__main__":
    main()

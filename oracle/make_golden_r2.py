"""Round-2 fixtures from the UNMODIFIED reference code/model.py (imported from /root/reference/code; build container only --
the outputs are committed because /root/reference does not exist on the GPU box).  TEST INFRASTRUCTURE.

    python oracle/make_golden_r2.py

  unet_B.npz          model.py:70-94 at the BENCHMARK shape (1,1,257,1034) (3 s @ 44.1 kHz, BASELINE-literal), checkpoint seed 7
  unet_ckpt_{3,11}.npz, unet_ckpt_default.npz
                      the same forward at (1,1,257,188) on other checkpoints: seeded_state_dict(3), (11) and the reference's own
                      default initialisation (torch.manual_seed(5); UNet()) with its identity BatchNorm statistics
  unet_init_seed5.npz per-key fingerprints of that default-initialised state_dict (the drop-in's constructor must reproduce it)
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
from oracle.make_golden import magnitude_like  # noqa: E402


def main():
    import model as ref_model      # /root/reference/code/model.py
    gold = os.path.join(ROOT, "tests", "golden")
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    def run(sd, shape, seed, name, ckpt):
        net = ref_model.UNet(in_channels=1, num_classes=1)
        net.load_state_dict(sd, strict=True)
        net.eval()
        x = magnitude_like(shape, seed).half().float()          # inputs stored as float16: half the fixture size, exact round trip
        with torch.no_grad():
            y = net(x)
        np.savez_compressed(os.path.join(gold, f"{name}.npz"), x=x.numpy().astype(np.float16), y=y.numpy(), ckpt=np.array(ckpt))
        print(name, tuple(y.shape), float(y.abs().max()))

    run(seeded_state_dict(7), (1, 1, 257, 1034), 31, "unet_B", "seeded_state_dict(7)")
    run(seeded_state_dict(3), (1, 1, 257, 188), 32, "unet_ckpt_3", "seeded_state_dict(3)")
    run(seeded_state_dict(11), (1, 1, 257, 188), 33, "unet_ckpt_11", "seeded_state_dict(11)")
    torch.manual_seed(5)
    net = ref_model.UNet()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    run(sd, (1, 1, 257, 188), 34, "unet_ckpt_default", "torch.manual_seed(5); reference UNet() default init")
    keys = list(sd)
    fp = np.array([[float(v.double().sum()), float(v.double().abs().sum()), float(v.reshape(-1)[0]), float(v.reshape(-1)[-1])] for v in sd.values()])
    np.savez_compressed(os.path.join(gold, "unet_init_seed5.npz"), keys=np.array(keys), fingerprint=fp)
    print("init fingerprint", fp.shape)


if __name__ == "__main__":
    main()

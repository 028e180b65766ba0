#!/usr/bin/env python
"""Where the training step's time goes: forward + loss, backward, optimizer captured as three CUDA graphs and replayed separately
(batch 16 x (1,256,64) by default).  usage: python scripts/prof_train_phases.py [batch]"""
import sys

import torch

sys.path.insert(0, ".")
from audiodenoiser_b200.model import UNet
from audiodenoiser_b200.training import TrainEngine

b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = UNet()
eng = TrainEngine(net, device=dev, lr=1e-4)
x = torch.rand(b, 1, 256, 64, device=dev); y = torch.rand(b, 1, 256, 64, device=dev)
for _ in range(2):
    eng.train_step(x, y)
torch.cuda.synchronize()


def timed(g, reps=20):
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g_f, g_b, g_o, g_all = (torch.cuda.CUDAGraph() for _ in range(4))
with torch.cuda.graph(g_f):
    eng.zero_grad()
    out = eng.forward(x)
    losses, d_pred = eng.loss_and_grad(out, y)
with torch.cuda.graph(g_b):
    eng.backward(d_pred)
# backward() consumed the saved activations: re-run forward eagerly so that the optimizer graph has something consistent to work on
with torch.cuda.graph(g_o):
    eng.optimizer_step()
with torch.cuda.graph(g_all):
    eng.train_step(x, y)
print(f"batch {b}: whole step {timed(g_all):.3f} ms; forward+loss {timed(g_f):.3f}, backward {timed(g_b):.3f}, optimizer+pack {timed(g_o):.3f}")
for n_side in ("0", "1", "2"):
    import os
    os.environ["ADN_WGRAD_STREAMS"] = n_side
    eng.overlap_wgrad = n_side != "0"
    g = torch.cuda.CUDAGraph()
    out2 = eng.forward(x); _l, dp = eng.loss_and_grad(out2, y)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        out2 = eng.forward(x); _l, dp = eng.loss_and_grad(out2, y)
        eng.backward(dp)
    print(f"  forward+loss+backward with {n_side} wgrad side streams: {timed(g):.3f} ms")

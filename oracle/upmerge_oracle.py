"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

CPU restatement (float64 torch) of the algebra behind csrc/conv_upm.cu: the reference's UpSampleLayer.forward up to the first ReLU
(/root/reference/code/model.py:41-49 followed by :11 of DoubleConvLayer)

    x1 = ConvTranspose2d(Cl, Cu, kernel_size=2, stride=2)(x1); x1 = F.pad(x1, [0, dW, 0, dH]); x = cat([x2, x1], 1); z = Conv2d(3x3, pad 1)(x)

computed WITHOUT forming the up-sampled tensor.  A stride-2 2x2 transposed conv does not overlap, up(Y, X) = Wt[:, :, Y&1, X&1]^T low(Y>>1, X>>1) + bt,
so for the output pixels of one parity class (py, px) = (Y&1, X&1) the 3x3 conv over `up` is a 2x2 conv over `low`:

    z(2y+py, 2x+px) = sum_{ky,kx} W3_skip[ky,kx] skip(2y+py+ky-1, 2x+px+kx-1)
                    + sum_{dy,dx in {0,1}} Weff[py,px][dy,dx] low(y + dy - (1-py), x + dx - (1-px))
                    + sum over the taps (ky,kx) whose up-sampled pixel lies inside [0,2Hl) x [0,2Wl) of  W3_up[ky,kx] bt
    Weff[py,px][dy,dx] = sum over the taps with floor((py+ky-1)/2) + (1-py) == dy (same for x) of  W3_up[ky,kx] . Wt[(py+ky-1)&1][(px+kx-1)&1]

Parity status: pinned to torch's own conv_transpose2d / pad / cat / conv2d (the ops model.py calls) in tests/test_oracle_upmerge.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def low_index(p: int, k: int) -> int:
    """Index (0 / 1) of the low-resolution row (column) that tap k of a class with parity p reads, relative to the origin y - (1 - p)."""
    return (p + k + 1) // 2 - p


def merged_weights(w3: torch.Tensor, wt: torch.Tensor, c0: int) -> torch.Tensor:
    """w3: (Co, c0 + Cu, 3, 3), wt: (Cl, Cu, 2, 2) -> Weff (2, 2, 2, 2, Co, Cl) indexed [py][px][dy][dx][co][cl]."""
    co, cu = w3.shape[0], w3.shape[1] - c0
    cl = wt.shape[0]
    w_up = w3[:, c0:].double()
    out = torch.zeros(2, 2, 2, 2, co, cl, dtype=torch.float64)
    for py in range(2):
        for px in range(2):
            for ky in range(3):
                for kx in range(3):
                    dy, dx = low_index(py, ky), low_index(px, kx)
                    qy, qx = (py + ky - 1) & 1, (px + kx - 1) & 1
                    out[py, px, dy, dx] += w_up[:, :, ky, kx] @ wt[:, :, qy, qx].double().t()
    return out


def up_conv_merged(skip: torch.Tensor, low: torch.Tensor, w3: torch.Tensor, b3: torch.Tensor, wt: torch.Tensor, bt: torch.Tensor) -> torch.Tensor:
    """The conv output z (N, Co, H, W) by parity classes; skip (N, c0, H, W), low (N, Cl, Hl, Wl) with H - 2 Hl, W - 2 Wl in {0, 1}."""
    n, c0, h, w = skip.shape
    hl, wl = low.shape[2:]
    co = w3.shape[0]
    weff = merged_weights(w3, wt, c0)
    z = F.conv2d(skip.double(), w3[:, :c0].double(), b3.double(), padding=1)          # skip half + conv bias: an ordinary conv
    lowp = F.pad(low.double(), [1, 1, 1, 1])                                           # low(-1), low(Hl) read as zero
    wb = torch.einsum("ocyx,c->yxo", w3[:, c0:].double(), bt.double())                 # per-tap ConvTranspose-bias term
    for py in range(2):
        for px in range(2):
            ys, xs = range(py, h, 2), range(px, w, 2)
            ny, nx = len(ys), len(xs)
            if ny == 0 or nx == 0:
                continue
            acc = torch.zeros(n, co, ny, nx, dtype=torch.float64)
            for dy in range(2):
                for dx in range(2):
                    oy, ox = dy - (1 - py) + 1, dx - (1 - px) + 1                      # +1: the zero border of lowp
                    patch = lowp[:, :, oy:oy + ny, ox:ox + nx]
                    if patch.shape[2] < ny or patch.shape[3] < nx:                    # odd H / W: the last class row / column reads past the border
                        patch = F.pad(patch, [0, nx - patch.shape[3], 0, ny - patch.shape[2]])
                    acc += torch.einsum("ol,nlyx->noyx", weff[py, px, dy, dx], patch)
            for iy, yy in enumerate(ys):                                               # bias terms of the taps that hit the up-sampled map
                for ix, xx in enumerate(xs):
                    for ky in range(3):
                        for kx in range(3):
                            if 0 <= yy + ky - 1 < 2 * hl and 0 <= xx + kx - 1 < 2 * wl:
                                acc[:, :, iy, ix] += wb[ky, kx]
            z[:, :, py::2, px::2] += acc
    return z


def up_conv_reference(skip, low, w3, b3, wt, bt):
    """The same through the ops model.py:41-49 calls."""
    h, w = skip.shape[2:]
    up = F.conv_transpose2d(low.double(), wt.double(), bt.double(), stride=2)
    up = F.pad(up, [0, w - up.shape[3], 0, h - up.shape[2]])
    return F.conv2d(torch.cat([skip.double(), up], 1), w3.double(), b3.double(), padding=1)

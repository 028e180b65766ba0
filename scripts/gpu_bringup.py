"""Bring-up diagnostics on a B200: runs every kernel once against the CPU oracle and prints error figures
(no asserts, so one call reports everything).  Usage: python scripts/gpu_bringup.py [stage ...]"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audiodenoiser_b200 import _lib, spectral, synth  # noqa: E402
from audiodenoiser_b200.checkpoint import seeded_state_dict  # noqa: E402
from oracle import stft_oracle, unet_oracle  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
S = lambda: _lib.stream_ptr()  # noqa: E731


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30)), float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def stage_stft():
    for L, center in ((16000, False), (24000, True), (24001, True), (512, False), (132300, True)):
        x = np.stack([synth.make_clip(i, "B")[:L] for i in range(3)])
        ref = np.stack([stft_oracle.stft_mag(xi.astype(np.float64), center) for xi in x])
        got = spectral.stft_mag_batched(torch.from_numpy(x).to(dev), center).cpu().numpy()
        print(f"stft L={L} center={center} shape={got.shape} max/norm rel = {rel(got, ref)}", flush=True)
    x = synth.make_clip(0, "R")
    refc = stft_oracle.stft(x.astype(np.float64), center=True)
    gotc = spectral.stft_complex_batched(torch.from_numpy(x).to(dev), True)[0].cpu().numpy()
    print("stft complex", rel(gotc.real, refc.real), rel(gotc.imag, refc.imag), flush=True)


def stage_istft():
    rng = np.random.default_rng(0)
    for T in (188, 2, 33, 1034):
        mag = np.abs(rng.standard_normal((2, 257, T))).astype(np.float32)
        ang = np.exp(2j * np.pi * rng.random((2, 257, T))).astype(np.complex64)
        ref = np.stack([stft_oracle.istft(mag[i].astype(np.float64) * ang[i].astype(np.complex128)) for i in range(2)])
        got = spectral.istft_batched(torch.from_numpy(mag).to(dev), torch.from_numpy(ang).to(dev)).cpu().numpy()
        print(f"istft T={T} shape={got.shape} rel={rel(got, ref)}", flush=True)
    x = synth.make_clip(1, "R")
    spec = spectral.stft_complex_batched(torch.from_numpy(x).to(dev), True)
    y = spectral.istft_batched(spec)[0].cpu().numpy()
    print("round trip", rel(y, x[: y.shape[0]]), flush=True)
    m = torch.from_numpy(np.abs(rng.standard_normal((1, 257, 40))).astype(np.float32)).to(dev)
    a = spectral.istft_batched(m, None, seed=1).cpu().numpy(); b = spectral.istft_batched(m, None, seed=1).cpu().numpy()
    c = spectral.istft_batched(m, None, seed=2).cpu().numpy()
    print("random phase: same seed equal", np.array_equal(a, b), "diff seed differs", not np.array_equal(a, c), "rms", float(np.sqrt((a ** 2).mean())), flush=True)


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def to_nhwc_bf16(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def from_nhwc(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def stage_conv():
    g = torch.Generator().manual_seed(0)
    cases = [(2, 64, 0, 64, 20, 19), (1, 64, 0, 128, 37, 26), (2, 128, 0, 256, 9, 6), (1, 256, 256, 256, 16, 11),
             (2, 64, 64, 64, 33, 28), (1, 512, 0, 1024, 4, 3), (3, 64, 0, 64, 257, 188)]
    for (n, c0, c1, co, h, w) in cases:
        ci = c0 + c1
        x0 = bf16_round(torch.randn(n, c0, h, w, generator=g))
        h1, w1 = (h - (h % 2), w - (w % 2)) if c1 else (0, 0)
        x1 = bf16_round(torch.randn(n, c1, h1, w1, generator=g)) if c1 else None
        wt = bf16_round(torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (9 * ci)) ** 0.5)
        scale = 0.5 + torch.rand(co, generator=g); shift = 0.1 * torch.randn(co, generator=g)
        xin = x0 if x1 is None else torch.cat([x0, F.pad(x1, [0, w - w1, 0, h - h1])], 1)
        ref = F.relu(F.conv2d(xin, wt, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
        wp = torch.empty((co, 9, ci), dtype=torch.bfloat16, device=dev)
        wd = wt.to(dev).contiguous()
        _lib.check(lib.adn_pack_conv3x3_weight_bf16(wd.data_ptr(), co, ci, wp.data_ptr(), S()))
        a0 = to_nhwc_bf16(x0).to(dev); a1 = to_nhwc_bf16(x1).to(dev) if c1 else None
        out = torch.zeros((n, h, w, co), dtype=torch.bfloat16, device=dev)
        pool = torch.zeros((n, h // 2, w // 2, co), dtype=torch.bfloat16, device=dev)
        sc, sh = scale.to(dev), shift.to(dev)
        t0 = time.time()
        st = lib.adn_conv3x3_bn_relu_bf16(a0.data_ptr(), c0, a1.data_ptr() if c1 else 0, c1, h1, w1, n, h, w, wp.data_ptr(), co,
                                          sc.data_ptr(), sh.data_ptr(), out.data_ptr(), pool.data_ptr(), S())
        torch.cuda.synchronize()
        got = from_nhwc(out.cpu())
        gp = from_nhwc(pool.cpu())
        print(f"conv n={n} c0={c0} c1={c1} co={co} {h}x{w}: status={st} rel={rel(got, ref)} pool={rel(gp, F.max_pool2d(bf16_round(ref), 2))} {time.time()-t0:.3f}s", flush=True)


def stage_convt():
    g = torch.Generator().manual_seed(1)
    for (n, ci, co, h, w) in [(2, 128, 64, 9, 7), (1, 1024, 512, 2, 3), (2, 256, 128, 16, 11)]:
        x = bf16_round(torch.randn(n, ci, h, w, generator=g))
        wt = bf16_round(torch.randn(ci, co, 2, 2, generator=g) * (1.0 / ci) ** 0.5)
        b = 0.1 * torch.randn(co, generator=g)
        ref = F.conv_transpose2d(x, wt, b, stride=2)
        wp = torch.empty((4, co, ci), dtype=torch.bfloat16, device=dev)
        wd = wt.to(dev).contiguous()
        _lib.check(lib.adn_pack_convt2x2_weight_bf16(wd.data_ptr(), ci, co, wp.data_ptr(), S()))
        a = to_nhwc_bf16(x).to(dev); bd = b.to(dev)
        out = torch.zeros((n, 2 * h, 2 * w, co), dtype=torch.bfloat16, device=dev)
        st = lib.adn_convt2x2_bf16(a.data_ptr(), ci, n, h, w, wp.data_ptr(), co, bd.data_ptr(), out.data_ptr(), S())
        torch.cuda.synchronize()
        print(f"convT n={n} ci={ci} co={co} {h}x{w}: status={st} rel={rel(from_nhwc(out.cpu()), ref)}", flush=True)


def stage_unet():
    from audiodenoiser_b200.model import UNet
    sd = seeded_state_dict(7)
    net = UNet().eval()
    net.load_state_dict(sd)
    for name in ("small", "train", "test"):
        z = np.load(os.path.join(ROOT, "tests", "golden", f"unet_{name}.npz"))
        x = torch.from_numpy(z["x"]).to(dev)
        t0 = time.time()
        y = net(x)
        torch.cuda.synchronize()
        print(f"unet {name}: rel(max,norm)={rel(y.cpu().numpy(), z['y'])} {time.time()-t0:.3f}s", flush=True)
    # per-level diagnostics against the oracle's intermediates
    z = np.load(os.path.join(ROOT, "tests", "golden", "unet_small.npz"))
    x = torch.from_numpy(z["x"])
    _, inter = unet_oracle.unet_forward(sd, x, return_intermediates=True)
    net(x.to(dev)); torch.cuda.synchronize()
    ws = list(net._ws.values())[0]
    names = {"down1": "s0", "down2": "s1", "down3": "s2", "down4": "s3", "bottle": "s4", "up1": "ub3", "up2": "ub2", "up3": "ub1"}
    for k, b in names.items():
        print(f"  level {k}: rel={rel(from_nhwc(ws[b].cpu()).numpy(), inter[k].numpy())}", flush=True)


STAGES = {"stft": stage_stft, "istft": stage_istft, "conv": stage_conv, "convt": stage_convt, "unet": stage_unet}

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "adn", lib.adn_version(), "device_check", lib.adn_device_check(), flush=True)
    for s in (sys.argv[1:] or list(STAGES)):
        print(f"=== {s}", flush=True)
        try:
            STAGES[s]()
        except Exception as e:  # noqa: BLE001
            print(f"STAGE {s} FAILED: {type(e).__name__}: {e}", flush=True)

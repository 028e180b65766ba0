#!/usr/bin/env python
"""STFT / iSTFT kernel throughput at BASELINE config 2 (4 096-clip dataset-creation batch) and at the bench batch.
Prints GB/s of algorithmic bytes against the measured HBM copy peak.  usage: python scripts/bench_spectral.py [reps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodenoiser_b200 import spectral  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
for name, n, length, center in (("c2_R_4096x24000", 4096, 24000, True), ("c2_train_4096x16000", 4096, 16000, False),
                                ("B_64x132300", 64, 132300, True), ("B_512x132300", 512, 132300, True)):
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand((n, length), device=dev, generator=g) * 2 - 1
    t = spectral.num_frames(length, center)
    mag = torch.empty((n, 257, t), device=dev)
    ms = timeit(lambda: spectral.stft_mag_batched(x, center, out=mag))
    b = n * (4 * length + 4 * 257 * t)
    out["stft_" + name] = {"ms": round(ms, 4), "GBps": round(b / ms / 1e6), "frac": round(b / ms / 1e6 / peak, 3)}
    if center:
        audio = torch.empty((n, 128 * (t - 1)), device=dev)
        ms = timeit(lambda: spectral.istft_batched(mag, None, seed=1, out=audio))
        b = n * (4 * 257 * t + 4 * 128 * (t - 1))
        out["istft_rand_" + name] = {"ms": round(ms, 4), "GBps": round(b / ms / 1e6), "frac": round(b / ms / 1e6 / peak, 3)}
        if n * 257 * t * 8 < 8e9:
            ph = torch.polar(torch.ones_like(mag), torch.rand_like(mag) * 6.2831853)
            ms = timeit(lambda: spectral.istft_batched(mag, ph, out=audio))
            b = n * (12 * 257 * t + 4 * 128 * (t - 1))
            out["istft_phasor_" + name] = {"ms": round(ms, 4), "GBps": round(b / ms / 1e6), "frac": round(b / ms / 1e6 / peak, 3)}
            del ph
    del x, mag
# noise mixing at the train-chunk shape: 12 algorithmic bytes per sample (clean + noise in, noisy out)
from audiodenoiser_b200 import noise as adn_noise  # noqa: E402
n, length = 4096, 16000
c = torch.rand((n, length), device=dev) - 0.5
z = torch.randn((n, length), device=dev)
o = torch.empty_like(c)
ms = timeit(lambda: adn_noise.mix_noise_snr_batched(c, z, 8.0, out=o))
b = n * length * 12
out["mix_noise_snr_4096x16000"] = {"ms": round(ms, 4), "GBps": round(b / ms / 1e6), "frac": round(b / ms / 1e6 / peak, 3)}
# resample front-end: 3 s clips at 44.1 kHz (mono and stereo) -> 8 kHz; 4 B per input sample and channel + 4 B out
from audiodenoiser_b200 import resample as adn_resample  # noqa: E402
for ch in (1, 2):
    n, length = 1024, 132300
    xin = torch.rand((n, ch, length), device=dev) - 0.5
    yo = torch.empty((n, 24000), device=dev)
    ms = timeit(lambda: adn_resample.resample_batched(xin, 44100, 8000, out=yo))
    b = n * (4 * ch * length + 4 * 24000)
    out[f"resample_44k1_to_8k_{n}x{ch}x{length}"] = {"ms": round(ms, 4), "GBps": round(b / ms / 1e6), "frac": round(b / ms / 1e6 / peak, 3),
                                                      "audio_s_per_s": round(n * 3.0 / (ms * 1e-3))}
    del xin, yo
for k, v in out.items():
    print(k, v)

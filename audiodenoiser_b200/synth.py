"""Deterministic synthetic IRMAS-shaped clips: a tone/music mixture plus white or
pink noise mixed at the reference's 8 dB with the reference's own formula
(code/create_train_dataset.py:148-157: rms with +1e-12, 10^(snr/20), sum, clip to [-1, 1]).

Host-side numpy only; used by bench.py and the tests to make inputs.  Two clip variants:
  (R) reference-faithful: 3 s @ 8 kHz  -> L = 24 000 -> T = 188 (center=True)
  (B) BASELINE-literal:   3 s @ 44.1 kHz -> L = 132 300 -> T = 1034
"""
from __future__ import annotations

import numpy as np

SNR_DB = 8.0                      # create_train_dataset.py:33
CLIP_SECONDS = 3.0
VARIANTS = {"R": (8000, 24000), "B": (44100, 132300)}


def _clean(rng: np.random.Generator, n: int, sr: int) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / sr
    k_parts = int(rng.integers(3, 7))
    f1 = float(np.exp(rng.uniform(np.log(80.0), np.log(1000.0))))
    y = np.zeros(n, dtype=np.float64)
    for k in range(1, k_parts + 1):
        fk = k * f1 * (1.0 + rng.uniform(-0.002, 0.002))
        if fk >= 0.45 * sr:
            break
        onset = rng.uniform(0.0, 0.5 * n / sr)
        decay = rng.uniform(0.5, 4.0)
        env = np.where(t >= onset, np.exp(-(t - onset) * decay), 0.0)
        y += (1.0 / k) * np.sin(2.0 * np.pi * fk * t + rng.uniform(0, 2 * np.pi)) * env
    peak = np.max(np.abs(y))
    return 0.5 * y / (peak + 1e-12)


def _pink(rng: np.random.Generator, n: int) -> np.ndarray:
    spec = np.fft.rfft(rng.standard_normal(n))
    f = np.arange(spec.shape[0], dtype=np.float64)
    f[0] = 1.0
    return np.fft.irfft(spec / np.sqrt(f), n=n)


def make_clip(index: int, variant: str = "R", return_clean: bool = False):
    """Clip ``index`` (seed 1234 + index), float32, length L of the variant."""
    sr, n = VARIANTS[variant]
    rng = np.random.default_rng(1234 + int(index))
    clean = _clean(rng, n, sr)
    noise = rng.standard_normal(n) if index % 2 == 0 else _pink(rng, n)
    rms_c = np.sqrt(np.mean(clean ** 2) + 1e-12)
    rms_n = np.sqrt(np.mean(noise ** 2) + 1e-12)
    noise = noise * (rms_c / (10.0 ** (SNR_DB / 20.0)) / rms_n)
    noisy = np.clip(clean + noise, -1.0, 1.0).astype(np.float32)
    if return_clean:
        return noisy, clean.astype(np.float32)
    return noisy


def make_clips(count: int, variant: str = "R", start: int = 0, unique: int | None = None) -> np.ndarray:
    """(count, L) float32.  ``unique`` bounds how many distinct clips are synthesised (the rest are
    cyclic repeats with a deterministic per-repeat gain) so very large batches stay cheap to build."""
    sr, n = VARIANTS[variant]
    uniq = count if unique is None else min(count, unique)
    base = np.stack([make_clip(start + i, variant) for i in range(uniq)])
    if uniq == count:
        return base
    out = np.empty((count, n), dtype=np.float32)
    for i in range(count):
        out[i] = base[i % uniq] * np.float32(1.0 - 0.25 * ((i // uniq) % 3) / 3.0)
    return out

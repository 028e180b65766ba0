"""fp32 CPU restatement of the reference loss (code/loss.py:6-95) without torchaudio.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Pinned: tests/test_oracle_loss.py checks it against tests/golden/loss_*.npz, produced by
oracle/make_golden.py from the reference's own ``loss.CombinedPerceptualLoss`` (which
goes through ``torchaudio.transforms.MelSpectrogram``).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

FFT_SIZES = (63, 32, 16)      # loss.py:7
HOP_LENGTHS = (16, 8, 4)
W_STFT, W_MEL, W_L1 = 0.4, 0.4, 0.2   # loss.py:79-81
MEL_SR, MEL_NFFT, MEL_HOP, MEL_NMELS = 8000, 63, 16, 64   # loss.py:38


def time_envelope(x: torch.Tensor) -> torch.Tensor:
    """loss.py:14-20 / 46-52: (B,1,F,T) -> mean over F -> (B,T)."""
    if x.dim() == 4:
        x = x.mean(dim=2)
    if x.dim() == 3 and x.size(1) == 1:
        x = x.squeeze(1)
    return x


def multiscale_stft_loss(pred, target):
    """MultiScaleSTFTLoss.forward, loss.py:12-35: rectangular-window STFT magnitudes of the
    time envelope at three scales, L1, averaged."""
    p, t = time_envelope(pred), time_envelope(target)
    loss = 0.0
    for n, hop in zip(FFT_SIZES, HOP_LENGTHS):
        w = torch.ones(n, dtype=p.dtype)
        pm = torch.abs(torch.stft(p, n_fft=n, hop_length=hop, return_complex=True, pad_mode="constant", window=w))
        tm = torch.abs(torch.stft(t, n_fft=n, hop_length=hop, return_complex=True, pad_mode="constant", window=w))
        loss = loss + F.l1_loss(pm, tm)
    return loss / len(FFT_SIZES)


def mel_filterbank(n_freqs=MEL_NFFT // 2 + 1, f_min=0.0, f_max=MEL_SR / 2, n_mels=MEL_NMELS, sr=MEL_SR):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') -> (n_freqs, n_mels)."""
    all_freqs = torch.linspace(0, sr // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


def mel_spectrogram(x: torch.Tensor) -> torch.Tensor:
    """torchaudio MelSpectrogram(sr=8000, n_fft=63, hop=16, n_mels=64) defaults: periodic Hann(63),
    power 2, center=True reflect padding, HTK mel, no norm.  x: (..., T) -> (..., 64, frames)."""
    w = torch.hann_window(MEL_NFFT, periodic=True, dtype=x.dtype)
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    s = torch.stft(x2, n_fft=MEL_NFFT, hop_length=MEL_HOP, win_length=MEL_NFFT, window=w, center=True,
                   pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    power = s.abs().pow(2.0)                                   # (B, 32, frames)
    mel = torch.matmul(power.transpose(-1, -2), mel_filterbank().to(x.dtype)).transpose(-1, -2)
    return mel.reshape(shape[:-1] + mel.shape[-2:])


def mel_loss(pred, target):
    """MelSpectrogramLoss.forward, loss.py:44-69 (per-sample loop -> stacked (B,1,64,frames))."""
    p, t = time_envelope(pred), time_envelope(target)
    return F.l1_loss(mel_spectrogram(p.unsqueeze(1)), mel_spectrogram(t.unsqueeze(1)))


def combined_loss(pred, target):
    """CombinedPerceptualLoss.forward, loss.py:83-95 -> (total, stft, mel, l1)."""
    s = multiscale_stft_loss(pred, target)
    m = mel_loss(pred, target)
    l1 = F.l1_loss(pred, target)
    return W_STFT * s + W_MEL * m + W_L1 * l1, s, m, l1

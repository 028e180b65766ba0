"""Drop-in for the reference's code/loss.py: ``CombinedPerceptualLoss`` (+ its two component losses), forward pass on the GPU.

Same constructor (no arguments), same attributes (``w_stft = 0.4``, ``w_mel = 0.4``, ``w_l1 = 0.2``), same call:
``total, stft, mel, l1 = criterion(pred, target)`` with (B, 1, F, T) float32 tensors, returning four 0-dim tensors on the
inputs' device (test.py:118-122 passes host tensors, train.py:68,85 CUDA tensors).  The arithmetic runs in the CUDA kernels of csrc/loss.cu through
``adn_combined_loss_f32``.  When ``pred`` requires grad the four values are part of the autograd graph: ``loss.backward()``
(train.py:69) runs ``adn_combined_loss_backward_f32``.  No CPU code path (host tensors are staged to the GPU and back).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib

_MEL_SR, _MEL_NFFT, _MEL_NMELS = 8000, 63, 64       # loss.py:38


def mel_filterbank() -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(n_freqs=32, f_min=0, f_max=4000, n_mels=64, sample_rate=8000, norm=None,
    mel_scale='htk') -> (32, 64) float32, the constant matrix inside MelSpectrogram (loss.py:40-42)."""
    n_freqs = _MEL_NFFT // 2 + 1
    all_freqs = torch.linspace(0, _MEL_SR // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + (_MEL_SR / 2) / 700.0)
    m_pts = torch.linspace(m_min, m_max, _MEL_NMELS + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0).float().contiguous()


class _LossKernel:
    """Per-device constants and scratch for adn_combined_loss_f32."""

    def __init__(self):
        self._fb = {}
        self._ws = {}

    def __call__(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        torch_ = _lib.require_cuda()
        if not (isinstance(pred, torch_.Tensor) and pred.is_cuda and target.is_cuda):
            raise _lib.AdnError("the loss kernels take CUDA tensors (host tensors are staged by CombinedPerceptualLoss.forward)")
        if pred.shape != target.shape:
            raise ValueError("pred and target must have the same shape")
        if pred.dim() == 3:                       # (B, F, T) is accepted as (B, 1, F, T)
            pred, target = pred.unsqueeze(1), target.unsqueeze(1)
        if pred.dim() != 4 or pred.shape[1] != 1:
            raise ValueError("expected (B, 1, F, T) tensors")
        pred = pred.detach().float().contiguous(); target = target.detach().float().contiguous()
        b, _, f, t = pred.shape
        dev = pred.device
        lib = _lib.load()
        if dev not in self._fb:
            self._fb[dev] = mel_filterbank().to(dev)
        need = int(lib.adn_loss_workspace_bytes(b, f, t))
        ws = self._ws.get(dev)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
            self._ws[dev] = ws
        out = torch.empty(4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = lib.adn_combined_loss_f32(pred.data_ptr(), target.data_ptr(), b, f, t, self._fb[dev].data_ptr(), ws.data_ptr(),
                                           out.data_ptr(), _lib.stream_ptr())
        _lib.check(st, "adn_combined_loss_f32")
        return out


_KERNEL = _LossKernel()


class _LossFunction(torch.autograd.Function):
    """(total, stft, mel, l1) = f(pred, target) as one autograd node (gradient with respect to pred only, like train.py needs)."""

    @staticmethod
    def forward(ctx, pred, target):
        ctx.save_for_backward(pred, target)
        return _KERNEL(pred, target)

    @staticmethod
    def backward(ctx, g4):
        pred, target = ctx.saved_tensors
        g = [float(v) for v in g4.detach().cpu().tolist()]
        c_stft, c_mel, c_l1 = 0.4 * g[0] + g[1], 0.4 * g[0] + g[2], 0.2 * g[0] + g[3]
        shape = pred.shape
        p = pred.detach().float().contiguous(); t = target.detach().float().contiguous()
        if p.dim() == 3:
            p, t = p.unsqueeze(1), t.unsqueeze(1)
        b, _, f, tt = p.shape
        lib = _lib.load()
        dev = p.device
        need = int(lib.adn_loss_backward_workspace_bytes(b, f, tt))
        ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
        d_pred = torch.empty_like(p)
        if dev not in _KERNEL._fb:
            _KERNEL._fb[dev] = mel_filterbank().to(dev)
        with torch.cuda.device(dev):
            st = lib.adn_combined_loss_backward_f32(p.data_ptr(), t.data_ptr(), b, f, tt, _KERNEL._fb[dev].data_ptr(), c_stft, c_mel, c_l1,
                                                    ws.data_ptr(), d_pred.data_ptr(), _lib.stream_ptr())
        _lib.check(st, "adn_combined_loss_backward_f32")
        return d_pred.reshape(shape), None


def _values(pred, target):
    """(total, stft, mel, l1) as a 4-vector on pred's device.  HOST tensors -- what test.py:119-121 passes -- are copied to the
    current GPU, reduced by the same kernels, and the four values are copied back (host in / host out, not a CPU code path)."""
    _lib.require_cuda()
    if not isinstance(pred, torch.Tensor) or not isinstance(target, torch.Tensor):
        raise TypeError("pred and target must be torch tensors")
    if not pred.is_cuda:
        dev = target.device if target.is_cuda else torch.device("cuda", torch.cuda.current_device())
        return _values(pred.to(dev), target.to(dev)).cpu()
    if not target.is_cuda:
        target = target.to(pred.device)
    if torch.is_grad_enabled() and pred.requires_grad:
        return _LossFunction.apply(pred, target)
    return _KERNEL(pred, target)


class MultiScaleSTFTLoss(nn.Module):
    """loss.py:6-35."""

    def __init__(self, fft_sizes=[63, 32, 16], hop_lengths=[16, 8, 4]):  # noqa: B006 - the reference's signature
        super().__init__()
        if list(fft_sizes) != [63, 32, 16] or list(hop_lengths) != [16, 8, 4]:
            raise ValueError("the B200 build is specialised to the reference's fft_sizes / hop_lengths")
        self.fft_sizes, self.hop_lengths = list(fft_sizes), list(hop_lengths)

    def forward(self, pred, target):
        return _values(pred, target)[1]


class MelSpectrogramLoss(nn.Module):
    """loss.py:37-69."""

    def __init__(self, sample_rate=8000, n_mels=64, n_fft=63, hop_length=16):
        super().__init__()
        if (sample_rate, n_mels, n_fft, hop_length) != (8000, 64, 63, 16):
            raise ValueError("the B200 build is specialised to the reference's MelSpectrogram(8000, 63, 16, 64)")

    def forward(self, pred, target):
        return _values(pred, target)[2]


class CombinedPerceptualLoss(nn.Module):
    """loss.py:71-95."""

    def __init__(self):
        super().__init__()
        self.stft_loss = MultiScaleSTFTLoss()
        self.mel_loss = MelSpectrogramLoss()
        self.l1_loss = nn.L1Loss()
        self.w_stft = 0.4
        self.w_mel = 0.4
        self.w_l1 = 0.2

    def values(self, pred, target):
        """The four values as ONE device (or host) 4-vector (total, stft, mel, l1) -- what forward() unpacks."""
        return _values(pred, target)

    def forward(self, pred, target):
        out = _values(pred, target)
        if (self.w_stft, self.w_mel, self.w_l1) != (0.4, 0.4, 0.2):      # weights edited after construction
            total = self.w_stft * out[1] + self.w_mel * out[2] + self.w_l1 * out[3]
        else:
            total = out[0]
        return total, out[1], out[2], out[3]

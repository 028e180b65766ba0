"""Noise mixing of the reference's add_noise (create_train_dataset.py:105-159) on the GPU, batched over clips:
the dataset-creation path clean chunk -> noisy chunk -> |STFT| without leaving the device.  torch is used only for device
memory and streams; no CPU fallback."""
from __future__ import annotations

import numpy as np

from . import _lib

SNR_DB = 8.0                                   # create_train_dataset.py:33
CANCEL_BLOCK, CANCEL_HALF, CANCEL_FACTOR = 16000, 8000, -0.8          # :127-132


def _pair(clean, other=None):
    torch = _lib.require_cuda()
    if not isinstance(clean, torch.Tensor) or not clean.is_cuda:
        raise _lib.AdnError("expected CUDA tensors (no CPU fallback)")
    if clean.dim() == 1:
        clean = clean.unsqueeze(0)
    clean = clean.float().contiguous()
    if other is not None:
        if other.dim() == 1:
            other = other.unsqueeze(0)
        if tuple(other.shape) != tuple(clean.shape):
            raise ValueError("noise must have the clean signal's shape")
        other = other.to(device=clean.device, dtype=torch.float32).contiguous()
    return torch, clean, other


def mix_noise_snr_batched(clean, noise, snr_db: float = SNR_DB, out=None):
    """(N, L) clean + (N, L) noise -> clip(clean + noise scaled to ``snr_db`` against each clean row, -1, 1)  (:148-157)."""
    torch, clean, noise = _pair(clean, noise)
    out = torch.empty_like(clean) if out is None else out
    with torch.cuda.device(clean.device):
        st = _lib.load().adn_mix_noise_snr_f32(clean.data_ptr(), noise.data_ptr(), clean.shape[0], clean.shape[1], float(snr_db),
                                               out.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "adn_mix_noise_snr_f32")
    return out


def mix_noise_cancel_batched(clean, block_flags, out=None):
    """(N, L) clean, (N, ceil(L/16000)) uint8 flags -> the "noise_cancellation" chunks (:123-135)."""
    torch, clean, _ = _pair(clean)
    flags = block_flags.to(device=clean.device, dtype=torch.uint8).contiguous()
    if flags.dim() == 1:
        flags = flags.unsqueeze(0)
    nb = -(-clean.shape[1] // CANCEL_BLOCK)
    if tuple(flags.shape) != (clean.shape[0], nb):
        raise ValueError(f"block_flags must be ({clean.shape[0]}, {nb})")
    out = torch.empty_like(clean) if out is None else out
    with torch.cuda.device(clean.device):
        st = _lib.load().adn_mix_noise_cancel_f32(clean.data_ptr(), flags.data_ptr(), clean.shape[0], clean.shape[1], CANCEL_BLOCK,
                                                  CANCEL_HALF, CANCEL_FACTOR, out.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "adn_mix_noise_cancel_f32")
    return out


def mix_noise_snr_host(clean: np.ndarray, noise: np.ndarray, snr_db: float = SNR_DB) -> np.ndarray:
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    c = torch.from_numpy(np.ascontiguousarray(clean, dtype=np.float32)).to(dev)
    n = torch.from_numpy(np.ascontiguousarray(noise, dtype=np.float32)).to(dev)
    return mix_noise_snr_batched(c, n, snr_db)[0].cpu().numpy()


def mix_noise_cancel_host(clean: np.ndarray, flags: np.ndarray) -> np.ndarray:
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    c = torch.from_numpy(np.ascontiguousarray(clean, dtype=np.float32)).to(dev)
    return mix_noise_cancel_batched(c, torch.from_numpy(np.ascontiguousarray(flags, dtype=np.uint8)))[0].cpu().numpy()

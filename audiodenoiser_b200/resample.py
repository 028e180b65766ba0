"""Mono mix-down + sample-rate conversion on the GPU: the numeric part of the reference's
``librosa.load(path, sr=SAMPLE_RATE)`` (create_train_dataset.py:204,225; create_test_dataset.py:139; test.py:80) after the file
has been decoded -- ``librosa.to_mono`` followed by ``librosa.resample`` to 8 kHz.

librosa 0.10 resamples with soxr_hq, a library absent from the reference tree and from this image; its filter is not specified by
the reference.  The kernel (csrc/resample.cu) implements the published polyphase algorithm of ``scipy.signal.resample_poly``
(Kaiser-windowed sinc, beta = 5, half length 10 * max(up, down), centre-aligned, ``ceil(L * up / down)`` output samples -- the same
output length librosa produces), so native-rate clips can flow into the 8 kHz path without leaving the device.  Parity: exact
against that algorithm (oracle/resample_oracle.py, float64), approximate against soxr (both are linear-phase low-passes at the
new Nyquist).  The filter design below is host logic in numpy; there is no CPU fallback for the convolution."""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np

from . import _lib

SAMPLE_RATE = 8000          # create_train_dataset.py:21, create_test_dataset.py:20, test.py:19


def rational_ratio(orig_sr: int, target_sr: int) -> tuple[int, int]:
    """(up, down) in lowest terms."""
    orig_sr, target_sr = int(orig_sr), int(target_sr)
    if orig_sr <= 0 or target_sr <= 0:
        raise ValueError("sample rates must be positive")
    g = math.gcd(orig_sr, target_sr)
    return target_sr // g, orig_sr // g


def output_length(len_in: int, up: int, down: int) -> int:
    """ceil(len_in * up / down): scipy.signal.resample_poly and librosa.resample agree on it."""
    return -(-int(len_in) * int(up) // int(down))


@lru_cache(maxsize=16)
def polyphase_filter(up: int, down: int, len_in: int):
    """The zero-padded, up-scaled prototype low-pass of scipy.signal.resample_poly(window=('kaiser', 5.0)) in polyphase order.

    Returns (table float32 (up, taps_pitch), taps, pre_remove): y[m] = sum_t x[j - t] * table[r, t] with c = (m + pre_remove) * down,
    j = c // up, r = c % up.  Depends on ``len_in`` only through scipy's trailing zero padding, which never changes a tap value."""
    max_rate = max(up, down)
    f_c = 1.0 / max_rate
    half_len = 10 * max_rate
    n = np.arange(-half_len, half_len + 1, dtype=np.float64)
    h = f_c * np.sinc(f_c * n) * np.kaiser(2 * half_len + 1, 5.0)      # firwin(2 * half_len + 1, f_c, window=('kaiser', 5.0)) ...
    h /= h.sum()                                                        # ... scaled to unit gain at DC
    h *= up
    n_pre_pad = down - half_len % down
    pre_remove = (half_len + n_pre_pad) // down
    g = np.concatenate([np.zeros(n_pre_pad), h])
    taps = -(-g.size // up)
    g = np.concatenate([g, np.zeros(taps * up - g.size)])
    pitch = -(-taps // 4) * 4                                           # rows read as float4: a multiple of 4 taps, zero-padded ...
    if (pitch // 4) % 2 == 0:
        pitch += 4                                                      # ... with an odd number of float4 per row (smem bank spread)
    table = np.zeros((up, pitch), np.float32)
    table[:, :taps] = g.reshape(taps, up).T.astype(np.float32)
    return table, taps, int(pre_remove)


_device_tables: dict = {}


def _table_on(device, up, down, len_in):
    torch = _lib.require_cuda()
    key = (str(device), up, down)
    hit = _device_tables.get(key)
    if hit is None:
        table, taps, pre = polyphase_filter(up, down, 0)
        hit = (torch.from_numpy(table).to(device), taps, pre)
        _device_tables[key] = hit
    return hit


def resample_batched(x, orig_sr: int, target_sr: int = SAMPLE_RATE, out=None):
    """CUDA float32 (N, L) mono or (N, C, L) planar multi-channel -> (N, ceil(L * target_sr / orig_sr)) mono at ``target_sr``."""
    torch = _lib.require_cuda()
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise _lib.AdnError("expected a CUDA tensor (no CPU fallback)")
    if x.dim() == 1:
        x = x.unsqueeze(0)
    if x.dim() == 2:
        x = x.unsqueeze(1)
    if x.dim() != 3 or x.shape[2] == 0:
        raise ValueError("expected (N, L) or (N, C, L) with L > 0")
    x = x.float().contiguous()
    n, c, length = x.shape
    up, down = rational_ratio(orig_sr, target_sr)
    if up == 1 and down == 1:
        return x.mean(dim=1) if c > 1 else x[:, 0].clone()
    len_out = output_length(length, up, down)
    table, taps, pre = _table_on(x.device, up, down, length)
    if out is None:
        out = torch.empty((n, len_out), dtype=torch.float32, device=x.device)
    elif tuple(out.shape) != (n, len_out) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float32 ({n}, {len_out}) tensor")
    with torch.cuda.device(x.device):
        st = _lib.load().adn_resample_poly_f32(x.data_ptr(), n, c, length, up, down, table.data_ptr(), taps, table.shape[1], pre,
                                               len_out, out.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "adn_resample_poly_f32")
    return out


def resample(y: np.ndarray, *, orig_sr: float, target_sr: float) -> np.ndarray:
    """Drop-in for ``librosa.resample(y, orig_sr=..., target_sr=...)`` on a mono (L,) or channel-first (C, L) array: resamples
    the last axis and keeps the channels (librosa semantics); float32 out."""
    torch = _lib.require_cuda()
    y = np.asarray(y, dtype=np.float32)
    if y.ndim not in (1, 2):
        raise ValueError("expected (L,) or (C, L)")
    dev = torch.device("cuda", torch.cuda.current_device())
    rows = torch.from_numpy(np.ascontiguousarray(y.reshape(-1, y.shape[-1]))).to(dev)
    out = resample_batched(rows, int(orig_sr), int(target_sr)).cpu().numpy()
    return out.reshape(y.shape[:-1] + (out.shape[-1],))


def load_decoded(audio: np.ndarray, native_sr: int, sr: int = SAMPLE_RATE, mono: bool = True) -> tuple[np.ndarray, int]:
    """What ``librosa.load(path, sr=sr, mono=mono)`` returns once soundfile has decoded ``path`` into ``audio`` -- (L,) or
    channel-first (C, L) float32 at ``native_sr``: mono mix, then resample.  Returns (audio, sr) like librosa.load."""
    torch = _lib.require_cuda()
    a = np.asarray(audio, dtype=np.float32)
    if a.ndim == 1 or not mono:
        return resample(a, orig_sr=native_sr, target_sr=sr), sr
    dev = torch.device("cuda", torch.cuda.current_device())
    x = torch.from_numpy(np.ascontiguousarray(a)).to(dev).unsqueeze(0)       # (1, C, L): mixed down inside the kernel
    return resample_batched(x, int(native_sr), int(sr))[0].cpu().numpy(), sr

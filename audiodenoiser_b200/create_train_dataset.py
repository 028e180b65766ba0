"""Drop-in for the hot-path function of the reference's code/create_train_dataset.py.

Same module constants and signature; numpy in, numpy out.  The arithmetic runs in the fused sm_100a STFT-magnitude
kernel through the C ABI (host buffers are copied in and out inside the call).  File walking, noise synthesis and
.npy writing stay with the caller (out of scope).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

SAMPLE_RATE = 8000          # create_train_dataset.py:21
FRAME_DURATION = 2.0        # :22
N_FFT = 512                 # :26
HOP_LENGTH = 128            # :27
SNR_DB = 8.0                # :33
NOISE_TYPES = ["white", "urban", "reverb", "noise_cancellation"]


def _stft_mag_host(audio_1d, center: bool) -> np.ndarray:
    audio = np.asarray(audio_1d)
    if audio.ndim != 1:
        raise ValueError("expected 1-D audio")
    out_dtype = np.float64 if audio.dtype == np.float64 else np.float32    # librosa: complex128 for float64 input
    x = np.ascontiguousarray(audio, dtype=np.float32)
    lib = _lib.load()
    t = lib.adn_stft_num_frames(x.shape[0], int(center))
    if t < 0:
        raise ValueError(f"Input signal length={x.shape[0]} is too small to analyze with n_fft={N_FFT}")   # librosa ParameterError
    _lib.require_cuda()
    mag = np.empty((N_FFT // 2 + 1, t), dtype=np.float32)
    st = lib.adn_stft_mag_host_f32(x.ctypes.data_as(ctypes.c_void_p), 1, x.shape[0], int(center),
                                   mag.ctypes.data_as(ctypes.c_void_p))
    _lib.check(st, "adn_stft_mag_host_f32")
    return mag if out_dtype == np.float32 else mag.astype(np.float64)


def audio_to_magnitude_spectrogram(audio_1d):
    """|STFT(n_fft=512, hop=128, center=False)|, shape (257, 1 + (L-512)//128) -- reference
    create_train_dataset.py:162-174.  float32 for float32 input (what the reference saves, :251-254); float64 input
    (white-noise chunks, :140) returns float64 values computed in float32."""
    return _stft_mag_host(audio_1d, center=False)

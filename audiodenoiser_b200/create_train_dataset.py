"""Drop-in for the hot-path function of the reference's code/create_train_dataset.py.

Same module constants and signatures; numpy in, numpy out.  The arithmetic runs in the sm_100a kernels through the C ABI
(host buffers are copied in and out inside the call): the fused STFT-magnitude kernel for ``audio_to_magnitude_spectrogram``
and the noise-mixing kernels for ``add_noise``.  The random draws of ``add_noise`` are made on the host with the
reference's own RNG calls in the reference's order, so seeding ``numpy.random`` / ``random`` reproduces the reference's chunks.
File walking, the pedalboard reverb and .npy writing stay with the caller (out of scope).
"""
from __future__ import annotations

import ctypes
import random

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


def match_audio_length(noise, target_len):
    """Noise of exactly ``target_len`` samples: tiled when shorter, a random snippet when longer (create_train_dataset.py:50-66;
    the snippet start is the reference's ``np.random.randint`` draw)."""
    n = len(noise)
    if n == target_len:
        return noise.copy()
    if n < target_len:
        return np.tile(noise, -(-target_len // n))[:target_len]
    start = np.random.randint(0, n - target_len)
    return noise[start:start + target_len]


def add_noise(clean_audio, noise_audio, noise_type, snr_db=SNR_DB):
    """create_train_dataset.py:105-159: "white" / "urban" noise at ``snr_db``, or "noise_cancellation"; result clipped to [-1, 1].
    The mixing (RMS, SNR scale, add, clip) runs on the GPU in float32; "white" returns float64 like the reference (its noise is
    ``np.random.randn``).  "reverb" needs pedalboard (JUCE) and is not part of the B200 path."""
    from . import noise as _noise
    clean = np.asarray(clean_audio)
    if clean.ndim != 1:
        raise ValueError("expected 1-D audio")
    n = len(clean)
    if noise_type == "reverb":
        raise NotImplementedError("the pedalboard reverb of create_train_dataset.py:87-102 is CPU glue outside the B200 hot path")
    if noise_type == "noise_cancellation":
        flags = np.array([random.random() < 0.8 for _ in range(0, n, 16000)], dtype=np.uint8)        # :128-129
        return _noise.mix_noise_cancel_host(clean, flags).astype(clean.dtype if clean.dtype.kind == "f" else np.float32)
    if noise_type == "white":
        noise = np.random.randn(n)                                                                    # :140
        out_dtype = np.float64
    else:
        if noise_audio is None or len(noise_audio) == 0:
            noise = np.zeros(n, dtype=np.float32)
        else:
            noise = match_audio_length(np.asarray(noise_audio), n)
        out_dtype = np.result_type(clean.dtype, noise.dtype) if clean.dtype.kind == "f" else np.float32
    return _noise.mix_noise_snr_host(clean, noise, snr_db).astype(out_dtype)

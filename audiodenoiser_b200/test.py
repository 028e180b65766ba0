"""Drop-in for the hot-path function of the reference's code/test.py: ``griffin_lim_reconstruction``.

What the reference's loop computes (test.py:29-48): a random unit phasor, then 50 x (istft -> stft) in which the target
magnitude is never re-imposed, then a final istft.  stft(istft(.)) is a projector, so the result equals ONE inverse STFT
of ``magnitude * angles`` (to ~1e-14 in float64).  The default path therefore runs the fused
iSTFT/overlap-add kernel once; ``faithful_loop=True`` runs the literal 50 round trips on the device instead.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

SAMPLE_RATE = 8000
N_FFT = 512
HOP_LENGTH_FFT = 128
NOISE_TYPES = ["white", "urban", "reverb", "noise_cancellation"]


def griffin_lim_reconstruction(magnitude_spectrogram, n_fft, hop_length, iterations=50, *, angles=None, rng=None,
                               faithful_loop=False):
    """magnitude (257, T) -> 1-D float64 audio of length hop*(T-1) (test.py:29-48).

    Keyword-only extras (not in the reference): ``angles`` -- complex unit phasors to use instead of a fresh random
    draw, so results are reproducible / comparable; ``rng`` -- a numpy Generator for the draw; ``faithful_loop``.
    """
    mag = np.asarray(magnitude_spectrogram)
    if mag.ndim != 2 or mag.shape[0] != N_FFT // 2 + 1 or n_fft != N_FFT or hop_length != HOP_LENGTH_FFT:
        raise ValueError("this build is specialised to the reference's n_fft=512, hop_length=128, (257, T) spectrograms")
    t = mag.shape[1]
    if angles is None:
        if rng is not None:
            angles = np.exp(2j * np.pi * rng.random(mag.shape))
        else:
            angles = np.exp(2j * np.pi * np.random.rand(*mag.shape))      # test.py:36 (global numpy RNG)
    angles = np.asarray(angles)
    if angles.shape != mag.shape:
        raise ValueError("angles must have the spectrogram's shape")
    _lib.require_cuda()
    if faithful_loop and iterations > 0:
        return _faithful(mag, angles, iterations)
    m32 = np.ascontiguousarray(mag, dtype=np.float32)
    a64 = np.ascontiguousarray(angles, dtype=np.complex64)
    out = np.empty(hop_length * (t - 1), dtype=np.float32)
    st = _lib.load().adn_istft_ola_host_f32(m32.ctypes.data_as(ctypes.c_void_p), a64.ctypes.data_as(ctypes.c_void_p), 0, 1, t,
                                            out.ctypes.data_as(ctypes.c_void_p))
    _lib.check(st, "adn_istft_ola_host_f32")
    return out.astype(np.float64)      # the reference returns float64 (complex128 angles, test.py:36-37)


def _faithful(mag, angles, iterations):
    import torch
    from . import spectral
    dev = torch.device("cuda", torch.cuda.current_device())
    spec = torch.from_numpy(np.ascontiguousarray(mag * angles, dtype=np.complex64)).to(dev).unsqueeze(0)
    for _ in range(iterations):
        audio = spectral.istft_batched(spec)
        spec = spectral.stft_complex_batched(audio, center=True)        # |S| * exp(i*angle(S)) == S (test.py:44-46)
    return spectral.istft_batched(spec)[0].double().cpu().numpy()

"""CPU restatement of the reference's noise synthesis: ``match_audio_length`` and ``add_noise``
(code/create_train_dataset.py:50-66 and :105-159; duplicated at code/create_test_dataset.py:43-133), minus the pedalboard reverb
branch (:116-121, JUCE, out of scope).  TEST INFRASTRUCTURE (see oracle/__init__.py).

The random draws are made with the same host RNG calls in the same order as the reference (``np.random.randn``,
``np.random.randint``, ``random.random``), so seeding the two global generators reproduces the reference bit for bit.

Pinned: tests/test_oracle_noise.py checks it against tests/golden/noise_small.npz, produced by oracle/make_golden_noise.py by
executing the reference's OWN function definitions (extracted from create_train_dataset.py with ``ast``, since the module itself
imports librosa / soundfile / pedalboard, which are absent here).
"""
from __future__ import annotations

import random

import numpy as np

SNR_DB = 8.0                 # create_train_dataset.py:33
BLOCK, HALF, FACTOR, P_CANCEL = 16000, 8000, -0.8, 0.8     # :127-132


def match_audio_length(noise, target_len):
    """create_train_dataset.py:50-66."""
    if len(noise) == target_len:
        return noise.copy()
    if len(noise) < target_len:
        return np.tile(noise, int(np.ceil(target_len / len(noise))))[:target_len]
    start = np.random.randint(0, len(noise) - target_len)
    return noise[start:start + target_len]


def cancellation_flags(clean_len):
    """The per-block draws of :128-129: one random.random() per 2 s block."""
    return np.array([random.random() < P_CANCEL for _ in range(0, clean_len, BLOCK)], dtype=np.uint8)


def add_noise(clean_audio, noise_audio, noise_type, snr_db=SNR_DB):
    """create_train_dataset.py:105-159 for "white", "urban", "noise_cancellation"."""
    clean_len = len(clean_audio)
    if noise_type == "reverb":
        raise NotImplementedError("pedalboard reverb (create_train_dataset.py:87-102) is out of scope")
    if noise_type == "noise_cancellation":
        noise = np.zeros_like(clean_audio)
        flags = cancellation_flags(clean_len)
        for b, i in enumerate(range(0, clean_len, BLOCK)):
            if flags[b]:
                half_end = min(i + HALF, min(i + BLOCK, clean_len))
                noise[i:half_end] = FACTOR * clean_audio[i:half_end]
        return np.clip(clean_audio + noise, -1.0, 1.0)
    if noise_type == "white":
        noise_audio = np.random.randn(clean_len)
    elif noise_audio is None or len(noise_audio) == 0:
        noise_audio = np.zeros(clean_len, dtype=np.float32)
    else:
        noise_audio = match_audio_length(noise_audio, clean_len)
    clean_rms = np.sqrt(np.mean(clean_audio ** 2) + 1e-12)
    noise_rms = np.sqrt(np.mean(noise_audio ** 2) + 1e-12)
    desired = clean_rms / (10.0 ** (snr_db / 20.0))
    if noise_rms > 1e-9:
        noise_audio = noise_audio * (desired / noise_rms)
    else:
        noise_audio = np.zeros_like(clean_audio)
    return np.clip(clean_audio + noise_audio, -1.0, 1.0)

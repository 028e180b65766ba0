"""Drop-in for the hot-path function of the reference's code/create_test_dataset.py (numpy in, numpy out)."""
from __future__ import annotations

from .create_train_dataset import _stft_mag_host

SAMPLE_RATE = 8000          # create_test_dataset.py:20
N_FFT = 512                 # :21
HOP_LENGTH_FFT = 128        # :22
SNR_DB = 8.0
NOISE_TYPES = ["white", "urban", "reverb", "noise_cancellation"]


def audio_to_spectrogram(audio):
    """|STFT(n_fft=512, hop=128)| with librosa's default center=True / zero padding, shape (257, 1 + L//128) --
    reference create_test_dataset.py:35-41."""
    return _stft_mag_host(audio, center=True)

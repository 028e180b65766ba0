"""float64 numpy restatement of the librosa-0.10.2 STFT / iSTFT arithmetic that the
reference executes on its hot path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED by the reference itself: librosa==0.10.2.post1 (reference
requirements.txt:10; scipy==1.14.1 :25) is a third-party dependency absent from
/root/reference and from this image.  What is restated is its published algorithm,
anchored on the reference's call sites:

  * ``librosa.stft(audio_1d, n_fft=512, hop_length=128, center=False)`` +
    ``librosa.magphase``            -- code/create_train_dataset.py:162-174
  * ``librosa.stft(audio, n_fft=512, hop_length=128)`` (center=True,
    pad_mode='constant' -- the 0.10 default) + ``magphase``
                                    -- code/create_test_dataset.py:35-41
  * ``librosa.istft(complex_spec, hop_length=128)`` and the 50x istft/stft loop
                                    -- code/test.py:29-48

It is cross-checked (tests/test_oracle_stft.py) against ``torch.stft`` /
``torch.istft`` in float64 and against analytic known answers.
"""
from __future__ import annotations

import numpy as np
import scipy.fft

N_FFT = 512
HOP = 128
N_BINS = N_FFT // 2 + 1


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True): periodic Hann, w[0] = 0."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def num_frames(length: int, center: bool, n_fft: int = N_FFT, hop: int = HOP) -> int:
    """Frame count of librosa.stft: 1 + (L_padded - n_fft)//hop."""
    padded = length + (n_fft if center else 0)
    if padded < n_fft:
        raise ValueError(f"input of length {length} is too short for n_fft={n_fft} with center={center}")
    return 1 + (padded - n_fft) // hop


def frame_starts(length: int, center: bool, n_fft: int = N_FFT, hop: int = HOP) -> np.ndarray:
    """Start offset (in UNPADDED sample coordinates, may be negative when centred)
    of every frame: frame t covers x[start_t : start_t + n_fft]."""
    t = np.arange(num_frames(length, center, n_fft, hop), dtype=np.int64)
    return t * hop - (n_fft // 2 if center else 0)


def stft(x: np.ndarray, n_fft: int = N_FFT, hop: int = HOP, center: bool = True) -> np.ndarray:
    """Complex STFT, shape (1 + n_fft//2, T).  float64 arithmetic; the result dtype
    follows librosa: complex64 for float32 input, complex128 for float64 input."""
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("expected 1-D audio")
    out_dtype = np.complex64 if x.dtype == np.float32 else np.complex128
    xf = x.astype(np.float64)
    if center:
        xf = np.pad(xf, (n_fft // 2, n_fft // 2), mode="constant")
    if xf.shape[0] < n_fft:
        raise ValueError(f"input of length {x.shape[0]} is too short for n_fft={n_fft}")
    t = 1 + (xf.shape[0] - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(t)[None, :]
    frames = xf[idx] * hann_periodic(n_fft)[:, None]
    d = scipy.fft.rfft(frames, n=n_fft, axis=0)
    return d.astype(out_dtype)


def stft_mag(x: np.ndarray, center: bool) -> np.ndarray:
    """|STFT| as the reference's audio_to_magnitude_spectrogram (center=False,
    create_train_dataset.py:162-174) / audio_to_spectrogram (center=True,
    create_test_dataset.py:35-41) return it."""
    return np.abs(stft(x, N_FFT, HOP, center))


def istft(d: np.ndarray, hop: int = HOP) -> np.ndarray:
    """librosa.istft(D, hop_length=hop) with every other argument default:
    n_fft = 2*(rows-1), Hann, center=True, length=None.  Returns
    hop*(T-1) samples, float64 for complex128 input / float32 for complex64."""
    d = np.asarray(d)
    n_fft = 2 * (d.shape[0] - 1)
    t = d.shape[1]
    out_dtype = np.float32 if d.dtype == np.complex64 else np.float64
    w = hann_periodic(n_fft)
    # c2r transform: imaginary parts of the DC and Nyquist rows are ignored
    y_frames = scipy.fft.irfft(d.astype(np.complex128), n=n_fft, axis=0) * w[:, None]
    full = n_fft + hop * (t - 1)
    y = np.zeros(full, dtype=np.float64)
    wss = np.zeros(full, dtype=np.float64)
    w2 = w * w
    for i in range(t):
        y[i * hop:i * hop + n_fft] += y_frames[:, i]
        wss[i * hop:i * hop + n_fft] += w2
    nz = wss > np.finfo(out_dtype).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2: full - n_fft // 2].astype(out_dtype)


def griffin_lim_reconstruction(mag: np.ndarray, n_fft: int, hop_length: int, iterations: int = 50,
                               angles: np.ndarray | None = None) -> np.ndarray:
    """Faithful restatement of code/test.py:29-48, including the loop that never
    re-imposes the target magnitude.  ``angles`` (complex unit phasors) is the only
    addition: the reference draws it from the unseeded global numpy RNG."""
    if angles is None:
        angles = np.exp(2j * np.pi * np.random.rand(*mag.shape))
    complex_spec = mag * angles
    for _ in range(iterations):
        audio = istft(complex_spec, hop=hop_length)
        new_spec = stft(audio, n_fft=n_fft, hop=hop_length, center=True)
        complex_spec = np.abs(new_spec) * np.exp(1j * np.angle(new_spec))
    return istft(complex_spec, hop=hop_length)

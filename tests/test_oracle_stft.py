"""The float64 STFT/iSTFT oracle against an independent implementation (torch.stft/istft, f64)
and the analytic known answers of SURVEY Appendix B.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import stft_oracle as so
from audiodenoiser_b200 import synth


@pytest.mark.parametrize("length,center,expect", [(16000, False, 122), (24000, True, 188), (132300, True, 1034),
                                                  (132300, False, 1030), (16000, True, 126), (512, False, 1), (0, True, 1)])
def test_frame_counts(length, center, expect):
    assert so.num_frames(length, center) == expect
    st = so.frame_starts(length, center)
    assert st[0] == (-256 if center else 0) and np.all(np.diff(st) == 128)


def test_too_short_raises():
    with pytest.raises(ValueError):
        so.stft(np.zeros(511, np.float32), center=False)


@pytest.mark.parametrize("center", [False, True])
def test_against_torch_stft(center):
    x = synth.make_clip(3, "R").astype(np.float64)[:16000]
    d = so.stft(x, center=center)
    ref = torch.stft(torch.from_numpy(x), 512, 128, window=torch.hann_window(512, periodic=True, dtype=torch.float64),
                     center=center, pad_mode="constant", return_complex=True).numpy()
    assert d.shape == ref.shape
    assert np.max(np.abs(d - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_impulse_frames():
    s = 5000
    x = np.zeros(24000); x[s] = 1.0
    m = so.stft_mag(x, center=True)
    hit = np.where(m.max(axis=0) > 0)[0]
    lo, hi = -(-(s + 256 - 511) // 128), (s + 256) // 128
    # the window is 0 at n=0, so a frame starting exactly on the impulse sees nothing
    expect = [t for t in range(lo, hi + 1) if (s + 256 - 128 * t) != 0]
    assert list(hit) == expect


def test_bin_centred_tone_and_dc():
    n = np.arange(16000)
    m = so.stft_mag(0.5 * np.cos(2 * np.pi * 32 * n / 512), center=False)
    assert np.allclose(m[32], 64.0, atol=1e-9) and np.allclose(m[31], 32.0, atol=1e-9) and np.allclose(m[33], 32.0, atol=1e-9)
    rest = np.delete(m, [31, 32, 33], axis=0)
    assert rest.max() < 1e-9
    m = so.stft_mag(np.full(16000, 0.25), center=False)
    assert np.allclose(m[0], 64.0) and np.allclose(m[1], 32.0) and m[2:].max() < 1e-9


def test_dtype_follows_librosa():
    assert so.stft(np.zeros(2048, np.float32)).dtype == np.complex64
    assert so.stft(np.zeros(2048, np.float64)).dtype == np.complex128
    assert so.stft_mag(np.zeros(2048, np.float32), True).dtype == np.float32


def test_round_trip_and_torch_istft():
    x = synth.make_clip(5, "R").astype(np.float64)
    d = so.stft(x, center=True)
    y = so.istft(d)
    assert y.shape == (128 * (d.shape[1] - 1),) and y.dtype == np.float64
    assert np.max(np.abs(y - x[: y.shape[0]])) < 1e-12
    ref = torch.istft(torch.from_numpy(d), 512, 128, window=torch.hann_window(512, periodic=True, dtype=torch.float64),
                      center=True).numpy()
    assert np.max(np.abs(y - ref)) < 1e-12


def test_istft_ignores_imag_of_dc_and_nyquist():
    rng = np.random.default_rng(0)
    d = rng.standard_normal((257, 9)) + 1j * rng.standard_normal((257, 9))
    d2 = d.copy(); d2[0] = d2[0].real; d2[256] = d2[256].real
    assert np.max(np.abs(so.istft(d) - so.istft(d2))) < 1e-14


def test_griffin_lim_loop_is_a_projector():
    """the reference loop never re-imposes the magnitude, so 50 iterations equal one iSTFT."""
    rng = np.random.default_rng(1)
    mag = np.abs(rng.standard_normal((257, 20))).astype(np.float32)
    ang = np.exp(2j * np.pi * rng.random((257, 20)))
    full = so.griffin_lim_reconstruction(mag, 512, 128, iterations=50, angles=ang)
    once = so.istft(mag * ang)
    assert full.shape == (128 * 19,) and full.dtype == np.float64
    assert np.max(np.abs(full - once)) <= 1e-12 * np.max(np.abs(once))

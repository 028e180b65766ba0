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


# ---------------------------------------------------------------------------------------------------------------------
# Pinning the unpinned oracle harder: three more implementations that share no code with the oracle
# (scipy.fft.rfft over a strided frame matrix) or with torch.stft: scipy's ShortTimeFFT class, scipy's legacy
# signal.stft / istft, and the defining O(N^2) DFT sum written out with a complex exponential matrix.  Plus a LIVE librosa
# comparison that runs wherever librosa is importable (it is absent from this image: requirements.txt:10 pins 0.10.2.post1).
def _clip(n=24000, seed=5):
    return synth.make_clip(seed, "R").astype(np.float64)[:n]


def test_against_scipy_shorttimefft():
    import scipy.signal as ss
    x = _clip()
    stf = ss.ShortTimeFFT(so.hann_periodic(512), hop=128, fs=1.0, fft_mode="onesided", mfft=512, scale_to=None, phase_shift=None)
    t = so.num_frames(len(x), True)
    ref = stf.stft(x, p0=0, p1=t)                       # slice p is centred on sample p*hop, zero padding outside: librosa center=True
    d = so.stft(x, center=True)
    assert ref.shape == d.shape == (257, 188)
    assert np.max(np.abs(d - ref)) <= 1e-12 * np.max(np.abs(ref))
    # inverse on an INCONSISTENT spectrogram (magnitude x random phase, what test.py:36-37 inverts): ShortTimeFFT synthesises with
    # the canonical dual window w / sum_shifts(w^2) = w / 1.5, librosa divides the overlap-add by the running window-sum of
    # squares -- identical wherever four frames overlap.  ShortTimeFFT.istft expects slices from p_min = -1 on: pad with zeros.
    rng = np.random.default_rng(3)
    c = np.abs(d) * np.exp(2j * np.pi * rng.random(d.shape))
    y = so.istft(c)
    s_full = np.concatenate([np.zeros((257, -stf.p_min)), c, np.zeros((257, 2))], axis=1)
    ref_y = stf.istft(s_full, k0=0, k1=128 * (t - 1))
    assert y.shape == ref_y.shape == (23936,)
    inner = slice(256, 23936 - 256)
    assert np.max(np.abs(y[inner] - ref_y[inner])) <= 1e-12 * np.max(np.abs(ref_y))


@pytest.mark.parametrize("n", [24000, 16000, 8100])
def test_against_scipy_legacy_stft(n):
    import scipy.signal as ss
    x = _clip(n)
    w = so.hann_periodic(512)
    _f, _t, z = ss.stft(x, window=w, nperseg=512, noverlap=384, nfft=512, boundary="zeros", padded=False, return_onesided=True,
                        scaling="spectrum")
    ref = z * w.sum()                                   # 'spectrum' scaling divides by sum(w)
    d = so.stft(x, center=True)
    assert ref.shape == d.shape
    assert np.max(np.abs(d - ref)) <= 1e-12 * np.max(np.abs(ref))
    # legacy istft (its own overlap-add + window-sum normalisation); it returns the full padded-trimmed signal, compare the
    # hop*(T-1) samples librosa.istft keeps
    _tt, yr = ss.istft(z, window=w, nperseg=512, noverlap=384, nfft=512, input_onesided=True, boundary=True, scaling="spectrum")
    y = so.istft(d)
    assert np.max(np.abs(y - yr[: len(y)])) <= 1e-12 * np.max(np.abs(y))


@pytest.mark.parametrize("center", [False, True])
def test_against_the_defining_dft_sum(center):
    """D[f, t] = sum_n w[n] x_t[n] exp(-2 pi i f n / 512) evaluated literally, no FFT anywhere."""
    x = _clip(4000, seed=9)
    xp = np.pad(x, (256, 256)) if center else x
    t = 1 + (len(xp) - 512) // 128
    n = np.arange(512)
    e = np.exp(-2j * np.pi * np.outer(np.arange(257), n) / 512.0)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * n / 512)
    ref = np.stack([e @ (w * xp[128 * i: 128 * i + 512]) for i in range(t)], axis=1)
    d = so.stft(x, center=center)
    assert d.shape == ref.shape
    assert np.max(np.abs(d - ref)) <= 1e-11 * np.max(np.abs(ref))
    # inverse by the literal sums: y_t[n] = (1/512) sum_k c_k Re(D[k,t] e^{+2 pi i k n/512}) (c2r weights 1,2,...,2,1; Im of DC /
    # Nyquist ignored), windowed overlap-add, division by the window sum of squares, trim 256 each side
    if center:
        c = np.full(257, 2.0); c[0] = c[256] = 1.0
        dd = d.astype(np.complex128).copy()
        dd[0] = dd[0].real; dd[256] = dd[256].real
        frames = ((c[:, None] * dd).T @ np.conj(e)).real / 512.0 * w        # (T, 512)
        full = 512 + 128 * (t - 1)
        acc, wss = np.zeros(full), np.zeros(full)
        for i in range(t):
            acc[128 * i: 128 * i + 512] += frames[i]
            wss[128 * i: 128 * i + 512] += w * w
        ref_y = (acc / np.where(wss > 0, wss, 1.0))[256: full - 256]
        y = so.istft(d)
        assert np.max(np.abs(y - ref_y)) <= 1e-11 * np.max(np.abs(ref_y))


def test_against_torchaudio_spectrogram():
    ta = pytest.importorskip("torchaudio")
    x = _clip()
    spec = ta.transforms.Spectrogram(n_fft=512, hop_length=128, power=1.0, center=True, pad_mode="constant",
                                     window_fn=lambda n: torch.hann_window(n, periodic=True, dtype=torch.float64))
    ref = spec(torch.from_numpy(x)).numpy()
    m = so.stft_mag(x, True)
    assert np.max(np.abs(m - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_against_live_librosa():
    """Runs wherever librosa is installed ("if the B200 box happens to have librosa, add a live comparison; never depend
    on it").  The calls are the reference's own: create_train_dataset.py:167-173, create_test_dataset.py:39-40, test.py:40."""
    librosa = pytest.importorskip("librosa")
    x32 = synth.make_clip(2, "R")
    for center in (False, True):
        d = librosa.stft(x32[:16000] if not center else x32, n_fft=512, hop_length=128, center=center)
        mag, _ = librosa.magphase(d)
        ours = so.stft(x32[:16000] if not center else x32, center=center)
        assert ours.dtype == d.dtype and ours.shape == d.shape
        assert np.max(np.abs(np.abs(ours) - mag)) <= 1e-6 * np.max(mag)
    d = so.stft(x32.astype(np.float64), center=True)
    assert np.max(np.abs(so.istft(d) - librosa.istft(d, hop_length=128))) <= 1e-12

"""Parity of the sm_100a STFT / iSTFT kernels (through the C ABI) against the float64 CPU oracle.

Tolerances (BASELINE.json north_star): frame indexing bit-exact; fp32 magnitudes within 1e-4 relative, defined as in
SURVEY 8d: max|d| / max|ref| per clip AND element-wise |d| <= 1e-4 * max(|ref|, 1e-3 * max|ref|)."""
import numpy as np
import pytest
import torch

from audiodenoiser_b200 import spectral, synth
from audiodenoiser_b200 import create_test_dataset, create_train_dataset, test as ref_test
from oracle import stft_oracle as so

pytestmark = pytest.mark.gpu
MAG_TOL = 1e-4


def dev():
    return torch.device("cuda", 0)


def assert_mag_close(got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    peak = np.max(np.abs(ref)) + 1e-30
    d = np.abs(got - ref)
    assert d.max() / peak <= MAG_TOL
    assert np.all(d <= MAG_TOL * np.maximum(np.abs(ref), 1e-3 * peak))


@pytest.mark.parametrize("length,center", [(16000, False), (24000, True), (24001, True), (512, False), (640, False), (1, True),
                                           (0, True), (300, True), (4095, True), (132300, True), (132300, False)])
def test_stft_mag_matches_oracle(length, center):
    x = np.stack([synth.make_clip(i, "B")[:length] for i in range(3)])
    ref = np.stack([so.stft_mag(xi.astype(np.float64), center) for xi in x])
    got = spectral.stft_mag_batched(torch.from_numpy(x).to(dev()), center).cpu().numpy()
    assert got.shape == (3, 257, so.num_frames(length, center))        # frame count bit-exact
    assert got.dtype == np.float32
    for g, r in zip(got, ref):
        assert_mag_close(g, r)


def test_frame_indexing_bit_exact_impulses():
    """An impulse lights exactly the frames whose window covers it."""
    length = 24000
    pos = [0, 1, 255, 256, 257, 5000, 12287, 12288, 23999]
    x = np.zeros((len(pos), length), np.float32)
    for i, p in enumerate(pos):
        x[i, p] = 1.0
    got = spectral.stft_mag_batched(torch.from_numpy(x).to(dev()), True).cpu().numpy()
    for i, p in enumerate(pos):
        ref = so.stft_mag(x[i].astype(np.float64), True)
        assert np.array_equal(np.where(got[i].max(axis=0) > 1e-6)[0], np.where(ref.max(axis=0) > 1e-6)[0])


def test_known_answers_tone_and_dc():
    n = np.arange(16000)
    x = np.stack([0.5 * np.cos(2 * np.pi * 32 * n / 512), np.full(16000, 0.25)]).astype(np.float32)
    m = spectral.stft_mag_batched(torch.from_numpy(x).to(dev()), False).cpu().numpy()
    assert np.allclose(m[0, 32], 64.0, rtol=1e-5) and np.allclose(m[0, 31], 32.0, rtol=1e-5) and np.allclose(m[0, 33], 32.0, rtol=1e-5)
    assert np.delete(m[0], [31, 32, 33], axis=0).max() < 64.0 * 1e-5
    assert np.allclose(m[1, 0], 64.0, rtol=1e-5) and np.allclose(m[1, 1], 32.0, rtol=1e-5) and m[1, 2:].max() < 64.0 * 1e-5


def test_linearity_and_ragged_stride():
    """|STFT(a x)| = |a| |STFT(x)|; rows taken from a wider buffer (clip_stride > length, unaligned rows)."""
    big = torch.from_numpy(np.stack([synth.make_clip(i, "R") for i in range(4)])).to(dev())
    view = big[:, 3:3 + 20001]                       # stride 24000, 12-byte offset: scalar load path
    a = spectral.stft_mag_batched(view, True)
    b = spectral.stft_mag_batched(view.contiguous() * -2.0, True)
    assert torch.allclose(b, 2.0 * a, rtol=1e-5, atol=1e-5)
    ref = so.stft_mag(view[2].cpu().numpy().astype(np.float64), True)
    assert_mag_close(a[2].cpu().numpy(), ref)


def test_stft_complex_matches_oracle():
    x = synth.make_clip(0, "R")
    ref = so.stft(x.astype(np.float64), center=True)
    got = spectral.stft_complex_batched(torch.from_numpy(x).to(dev()), True)[0].cpu().numpy()
    assert np.max(np.abs(got - ref)) <= 1e-5 * np.max(np.abs(ref))


@pytest.mark.parametrize("frames", [2, 3, 4, 5, 29, 30, 31, 33, 59, 188, 1034])
def test_istft_matches_oracle(frames):
    rng = np.random.default_rng(frames)
    mag = np.abs(rng.standard_normal((2, 257, frames))).astype(np.float32)
    ang = np.exp(2j * np.pi * rng.random((2, 257, frames))).astype(np.complex64)
    ref = np.stack([so.istft(mag[i].astype(np.float64) * ang[i].astype(np.complex128)) for i in range(2)])
    got = spectral.istft_batched(torch.from_numpy(mag).to(dev()), torch.from_numpy(ang).to(dev())).cpu().numpy()
    assert got.shape == (2, 128 * (frames - 1))
    assert np.max(np.abs(got - ref)) <= 1e-5 * np.max(np.abs(ref))


def test_istft_single_frame_is_empty():
    m = torch.ones(2, 257, 1, device=dev())
    assert spectral.istft_batched(m, torch.ones(2, 257, 1, dtype=torch.complex64, device=dev())).shape == (2, 0)


def test_round_trip_full_size():
    """istft(stft(x)) == x on the first 128*(T-1) samples -- size-independent property at the BASELINE clip shape."""
    x = np.stack([synth.make_clip(i, "B") for i in range(4)])
    xd = torch.from_numpy(x).to(dev())
    y = spectral.istft_batched(spectral.stft_complex_batched(xd, True))
    n = y.shape[1]
    assert n == 128 * 1033
    assert float((y - xd[:, :n]).abs().max()) <= 2e-6


def test_projector_property():
    """stft(istft(C)) applied twice equals once: the reference's 50-iteration loop is a no-op (test.py:39-46)."""
    rng = np.random.default_rng(3)
    c = torch.from_numpy((rng.standard_normal((1, 257, 64)) + 1j * rng.standard_normal((1, 257, 64))).astype(np.complex64)).to(dev())
    once = spectral.stft_complex_batched(spectral.istft_batched(c), True)
    twice = spectral.stft_complex_batched(spectral.istft_batched(once), True)
    assert float((once - twice).abs().max()) <= 1e-5 * float(once.abs().max())


def test_random_phase_is_seeded_and_deterministic():
    m = torch.rand(1, 257, 40, device=dev())
    a = spectral.istft_batched(m, None, seed=1); b = spectral.istft_batched(m, None, seed=1); c = spectral.istft_batched(m, None, seed=2)
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_random_phase_mode_equals_explicit_phasor():
    """The in-kernel phase generator is exported (adn_random_phasor_c64): seeded mode == explicit-phasor mode (to rounding: the
    seeded mode runs the warp-specialised kernel, the explicit-phasor modes the barrier-phased one -- same arithmetic, different
    fused-multiply-add contraction), the phasor has unit modulus and uniform phase, and the result matches the oracle istft of
    mag * phasor."""
    rng = np.random.default_rng(5)
    mag = torch.from_numpy(np.abs(rng.standard_normal((3, 257, 70))).astype(np.float32)).to(dev())
    ph = spectral.random_phasor(42, 3, 70)
    assert float((ph.abs() - 1).abs().max()) < 1e-5
    ang = torch.angle(ph).cpu().numpy().ravel()
    hist = np.histogram(ang, bins=16, range=(-np.pi, np.pi))[0]
    assert hist.min() > 0.8 * ang.size / 16 and abs(np.mean(np.exp(1j * ang))) < 0.02
    assert not torch.equal(ph[0], ph[1])                                   # clips get distinct phases
    a = spectral.istft_batched(mag, None, seed=42)
    b = spectral.istft_batched(mag, ph)
    assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max())
    ref = so.istft(mag[1].cpu().numpy().astype(np.float64) * ph[1].cpu().numpy().astype(np.complex128))
    assert np.max(np.abs(a[1].cpu().numpy() - ref)) <= 1e-5 * np.max(np.abs(ref))


# ------------------------------------------------------------------ the reference-named drop-in functions (numpy in/out)
def test_dropin_audio_to_magnitude_spectrogram():
    x = synth.make_clip(7, "R")[:16000]
    got = create_train_dataset.audio_to_magnitude_spectrogram(x)
    assert got.shape == (257, 122) and got.dtype == np.float32
    assert_mag_close(got, so.stft_mag(x.astype(np.float64), False))
    assert create_train_dataset.audio_to_magnitude_spectrogram(x.astype(np.float64)).dtype == np.float64
    with pytest.raises(ValueError):
        create_train_dataset.audio_to_magnitude_spectrogram(x[:300])


def test_dropin_audio_to_spectrogram():
    x = synth.make_clip(8, "R")
    got = create_test_dataset.audio_to_spectrogram(x)
    assert got.shape == (257, 188)
    assert_mag_close(got, so.stft_mag(x.astype(np.float64), True))


def test_dropin_griffin_lim_reconstruction():
    rng = np.random.default_rng(0)
    mag = np.abs(rng.standard_normal((257, 188))).astype(np.float32)
    ang = np.exp(2j * np.pi * rng.random((257, 188)))
    ref = so.griffin_lim_reconstruction(mag, 512, 128, iterations=50, angles=ang)      # the literal 50-iteration loop
    got = ref_test.griffin_lim_reconstruction(mag, 512, 128, angles=ang)
    assert got.shape == (23936,) and got.dtype == np.float64
    assert np.max(np.abs(got - ref)) <= 1e-5 * np.max(np.abs(ref))
    loop = ref_test.griffin_lim_reconstruction(mag, 512, 128, iterations=50, angles=ang, faithful_loop=True)
    assert np.max(np.abs(loop - ref)) <= 1e-4 * np.max(np.abs(ref))
    # default behaviour: fresh random phase each call
    a = ref_test.griffin_lim_reconstruction(mag, 512, 128); b = ref_test.griffin_lim_reconstruction(mag, 512, 128)
    assert a.shape == (23936,) and not np.array_equal(a, b)

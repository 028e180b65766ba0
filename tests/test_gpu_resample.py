"""Resample front-end on the GPU against the float64 scipy.signal.resample_poly oracle: the numeric tail of
librosa.load(path, sr=8000) -- mono mix + polyphase resampling (the reference's soxr filter itself is unpinned, see the oracle).
Tolerance: fp32 accumulation of <= 221 taps against float64 -> 2e-6 of full scale."""
import numpy as np
import pytest
import torch

from audiodenoiser_b200 import resample as rs
from audiodenoiser_b200 import spectral
from oracle import resample_oracle, stft_oracle

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
TOL = 2e-6


@pytest.mark.parametrize("orig,target,n,c,length", [(44100, 8000, 3, 1, 132300), (44100, 8000, 2, 2, 176400), (22050, 8000, 4, 1, 9999),
                                                    (16000, 8000, 1, 1, 513), (48000, 8000, 2, 3, 20001), (8000, 16000, 2, 1, 4000),
                                                    (44100, 8000, 1, 1, 1), (11025, 8000, 5, 1, 300)])
def test_resample_matches_oracle(orig, target, n, c, length):
    rng = np.random.default_rng(length + c)
    x = (0.5 * rng.standard_normal((n, c, length))).astype(np.float32)
    got = rs.resample_batched(torch.from_numpy(x).to(DEV), orig, target).cpu().numpy()
    up, down = rs.rational_ratio(orig, target)
    assert got.shape == (n, rs.output_length(length, up, down))
    for i in range(n):
        ref = resample_oracle.load_decoded(x[i] if c > 1 else x[i, 0], orig, target)
        assert np.max(np.abs(got[i] - ref)) <= TOL * max(1.0, float(np.max(np.abs(ref))))


def test_same_rate_is_a_mono_mix_and_numpy_entry_points():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 2, 1000)).astype(np.float32)
    y = rs.resample_batched(torch.from_numpy(x).to(DEV), 8000, 8000).cpu().numpy()
    assert np.allclose(y, x.mean(axis=1), atol=1e-7)
    stereo = rng.standard_normal((2, 44100)).astype(np.float32) * 0.3
    a, sr = rs.load_decoded(stereo, 44100)                               # librosa.load(path, sr=8000) after decoding
    assert sr == 8000 and a.shape == (8000,) and a.dtype == np.float32
    assert np.max(np.abs(a - resample_oracle.load_decoded(stereo, 44100, 8000))) <= TOL
    keep = rs.resample(stereo, orig_sr=44100, target_sr=8000)           # librosa.resample keeps the channels
    assert keep.shape == (2, 8000)
    assert np.max(np.abs(keep[1] - resample_oracle.resample(stereo[1], 44100, 8000))) <= TOL
    with pytest.raises(ValueError):
        rs.resample_batched(torch.zeros((1, 1, 0), device=DEV), 44100, 8000)


def test_native_rate_clip_flows_into_the_8khz_spectrogram():
    """44.1 kHz tone -> resample -> |STFT| on the device == oracle resample -> oracle STFT: the variant-B input reaches the
    variant-R shape (257, 188) the reference's scripts work on without leaving the GPU."""
    t = np.arange(3 * 44100) / 44100.0
    x = (0.4 * np.sin(2 * np.pi * 440.0 * t) + 0.1 * np.sin(2 * np.pi * 3000.0 * t)).astype(np.float32)
    y = rs.resample_batched(torch.from_numpy(x).to(DEV).unsqueeze(0), 44100, 8000)
    assert y.shape == (1, 24000)
    mag = spectral.stft_mag_batched(y, True).cpu().numpy()[0]
    ref = stft_oracle.stft_mag(resample_oracle.resample(x, 44100, 8000).astype(np.float32), center=True)
    assert mag.shape == ref.shape == (257, 188)
    assert np.max(np.abs(mag - ref)) <= 1e-4 * np.max(ref)

"""Host logic of the resample front-end without a GPU: the polyphase table + alignment that
audiodenoiser_b200/resample.py hands to the kernel, evaluated with a numpy restatement of the kernel's formula, must reproduce
scipy.signal.resample_poly (the oracle; the reference's own soxr resampler is an absent third-party library)."""
import numpy as np
import pytest

from audiodenoiser_b200 import resample as rs
from oracle import resample_oracle


def kernel_formula(x, up, down):
    """y[m] = sum_t x[j - t] * table[r, t], c = (m + pre_remove) * down, j = c // up, r = c % up (csrc/resample.cu), in float64."""
    table, taps, pre = rs.polyphase_filter(up, down, 0)
    t64 = table.astype(np.float64)
    n_out = rs.output_length(x.size, up, down)
    y = np.zeros(n_out)
    xp = np.concatenate([np.zeros(taps), x.astype(np.float64), np.zeros(taps + down)])
    for m in range(n_out):
        c = (m + pre) * down
        j, r = divmod(c, up)
        seg = xp[j + taps - np.arange(taps)]            # x[j - t], zero outside [0, L)
        y[m] = float(np.dot(seg, t64[r, :taps]))
    return y


@pytest.mark.parametrize("orig,target,length", [(44100, 8000, 5000), (22050, 8000, 3001), (16000, 8000, 1200), (8000, 16000, 700),
                                                (48000, 8000, 4100), (11025, 8000, 2000)])
def test_polyphase_table_reproduces_resample_poly(orig, target, length):
    rng = np.random.default_rng(orig + length)
    x = rng.standard_normal(length).astype(np.float32)
    up, down = rs.rational_ratio(orig, target)
    ref = resample_oracle.resample(x, orig, target)
    got = kernel_formula(x, up, down)
    assert got.shape == ref.shape == (rs.output_length(length, up, down),)
    assert np.max(np.abs(got - ref)) <= 2e-6 * np.max(np.abs(ref))      # the table is stored in float32


def test_ratio_and_length_helpers():
    assert rs.rational_ratio(44100, 8000) == (80, 441)
    assert rs.rational_ratio(16000, 8000) == (1, 2)
    assert rs.output_length(132300, 80, 441) == 24000                  # 3 s @ 44.1 kHz -> 3 s @ 8 kHz
    assert rs.output_length(132301, 80, 441) == 24001
    with pytest.raises(ValueError):
        rs.rational_ratio(0, 8000)


def test_oracle_mono_mix_and_dc_gain():
    stereo = np.stack([np.full(4410, 0.25, np.float32), np.full(4410, 0.75, np.float32)])
    y = resample_oracle.load_decoded(stereo, 44100, 8000)
    assert y.shape == (800,)
    assert np.allclose(y[100:700], 0.5, atol=1e-4)                     # unit DC gain away from the edges

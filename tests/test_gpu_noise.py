"""The noise-mixing kernels against the oracle restatement of the reference's add_noise, which is pinned bit for
bit to the reference's own function (tests/test_oracle_noise.py).  float32 mixing: 2e-6 of full scale."""
import random

import numpy as np
import pytest
import torch

from audiodenoiser_b200 import create_train_dataset, noise as adn_noise, spectral
from oracle import noise_oracle
from oracle.make_golden_noise import cases

pytestmark = pytest.mark.gpu
TOL = 2e-6


def test_dropin_add_noise_matches_reference_with_seeded_rng():
    """Same seeds -> same host draws (the reference's RNG calls in the reference's order) -> same noisy chunk, dtype included."""
    for name, clean, noise, nt, seed in cases():
        np.random.seed(seed); random.seed(seed)
        ref = noise_oracle.add_noise(clean.copy(), None if noise is None else noise.copy(), nt)
        np.random.seed(seed); random.seed(seed)
        got = create_train_dataset.add_noise(clean.copy(), None if noise is None else noise.copy(), nt)
        assert got.dtype == ref.dtype and got.shape == ref.shape, name
        assert np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64))) <= TOL, name
        assert got.min() >= -1.0 and got.max() <= 1.0
    with pytest.raises(NotImplementedError):
        create_train_dataset.add_noise(np.zeros(16000, np.float32), None, "reverb")


def test_batched_mix_and_fused_dataset_path():
    """(N, L) batches on the device: every row equals the per-clip oracle, and noisy -> |STFT| stays on the GPU."""
    rng = np.random.default_rng(1)
    n, length = 5, 16000
    clean = (rng.standard_normal((n, length)) * 0.2).astype(np.float32)
    clean[3] = 0.0                                                     # silent clean chunk: rms floor path
    noise = rng.standard_normal((n, length)).astype(np.float32)
    noise[4] = 0.0                                                     # silent noise: dropped (:152-155)
    d = torch.device("cuda", 0)
    got = adn_noise.mix_noise_snr_batched(torch.from_numpy(clean).to(d), torch.from_numpy(noise).to(d), 8.0)
    for i in range(n):
        c, z = clean[i], noise[i]
        crms, nrms = np.sqrt(np.mean(c.astype(np.float64) ** 2) + 1e-12), np.sqrt(np.mean(z.astype(np.float64) ** 2) + 1e-12)
        ref = np.clip(c + (z * (crms / 10 ** 0.4 / nrms) if nrms > 1e-9 else 0.0), -1, 1)
        assert np.max(np.abs(got[i].cpu().numpy() - ref)) <= TOL
    flags = torch.tensor([[1], [0], [1], [1], [0]], dtype=torch.uint8)
    nc = adn_noise.mix_noise_cancel_batched(torch.from_numpy(clean).to(d), flags).cpu().numpy()
    for i in range(n):
        ref = clean[i].copy()
        if flags[i, 0]:
            ref[:8000] = ref[:8000] + np.float32(-0.8) * ref[:8000]
        assert np.max(np.abs(nc[i] - np.clip(ref, -1, 1))) <= TOL
    mag = spectral.stft_mag_batched(got, center=False)
    assert mag.shape == (n, 257, 122) and torch.isfinite(mag).all()

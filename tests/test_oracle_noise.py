"""oracle/noise_oracle.py against the fixture produced by the reference's own add_noise (oracle/make_golden_noise.py).  CPU only."""
import os
import random

import numpy as np

from oracle import noise_oracle
from oracle.make_golden_noise import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden", "noise_small.npz")


def test_noise_oracle_matches_reference_bit_for_bit():
    gold = np.load(GOLD)
    for name, clean, noise, nt, seed in cases():
        np.random.seed(seed); random.seed(seed)
        y = noise_oracle.add_noise(clean.copy(), None if noise is None else noise.copy(), nt)
        assert y.dtype == gold[name].dtype and y.shape == gold[name].shape, name
        assert np.array_equal(y, gold[name]), name
        assert y.min() >= -1.0 and y.max() <= 1.0

"""SpectrogramDataset inside DataLoader WORKER processes (train.py:118-119: num_workers=4, pin_memory=True) -- the host-only half.
Workers have no CUDA context and this build has no CPU implementation of the transform's arithmetic: they only crop / zero-pad
(data_loader.py:54-72, no arithmetic); the float16 round trip (data_loader.py:41-42) is applied on the device by the consumer
(tests/test_gpu_data_loader.py::test_dataloader_with_worker_processes).  CPU only: the workers never touch CUDA."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from audiodenoiser_b200.data_loader import SpectrogramDataset


def _layout(a, target=(256, 64)):
    out = np.zeros((1,) + target, np.float32)
    f, t = min(target[0], a.shape[0]), min(target[1], a.shape[1])
    out[0, :f, :t] = a[:f, :t]
    return out


def test_workers_do_layout_only(tmp_path):
    rng = np.random.default_rng(2)
    shapes = [(257, 122), (257, 30), (100, 64), (300, 70)]
    for k, shp in enumerate(shapes):
        np.save(tmp_path / f"noisy_white_chunk_{k}.npy", (np.abs(rng.standard_normal(shp)) * 7.123).astype(np.float32))
        np.save(tmp_path / f"clean_white_chunk_{k}.npy", np.asfortranarray((np.abs(rng.standard_normal(shp)) * 0.3).astype(np.float32)))
    ds = SpectrogramDataset(str(tmp_path))
    got = list(DataLoader(ds, batch_size=2, shuffle=False, num_workers=2))
    assert len(got) == 2
    for b, (noisy, clean) in enumerate(got):
        assert noisy.shape == clean.shape == (2, 1, 256, 64) and noisy.dtype == torch.float32
        for j in range(2):
            k = 2 * b + j
            assert np.array_equal(noisy[j].numpy(), _layout(np.load(tmp_path / f"noisy_white_chunk_{k}.npy")))      # un-rounded: exact float32
            assert np.array_equal(clean[j].numpy(), _layout(np.load(tmp_path / f"clean_white_chunk_{k}.npy")))


def test_strict_workers_raise(tmp_path):
    np.save(tmp_path / "noisy_a.npy", np.ones((257, 122), np.float32))
    np.save(tmp_path / "clean_a.npy", np.ones((257, 122), np.float32))
    ds = SpectrogramDataset(str(tmp_path), strict_workers=True)
    with pytest.raises(RuntimeError, match="num_workers=0"):
        next(iter(DataLoader(ds, batch_size=1, num_workers=1)))

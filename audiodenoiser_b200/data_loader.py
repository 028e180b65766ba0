"""Drop-in for the reference's code/data_loader.py ``SpectrogramDataset`` (data_loader.py:7-72).

Same constructor, pairing rule (sorted ``clean*.npy`` / ``noisy*.npy`` in one folder, count assertion), ``__len__`` and
``__getitem__ -> (noisy, clean)`` each (1, 256, 64) float32.  The numeric transform of data_loader.py:41-72 --
float32 -> float16 round trip, zero-pad / crop to ``target_size`` -- runs in the CUDA kernel ``adn_spec_f16_crop_f32``;
file listing and ``np.load`` stay on the host (I/O glue).  ``load_batch`` is the batched device-resident fast path.
No CPU fallback: without a CUDA device the transform raises.

``DataLoader(dataset, num_workers=4, pin_memory=True)`` (train.py:118-119): worker PROCESSES cannot launch kernels (CUDA
cannot be re-initialised in a forked child, and one context per worker would be wasteful), and this build keeps no host
implementation of the arithmetic.  Inside a worker ``__getitem__`` therefore only does the layout part of
data_loader.py:54-72 -- crop / zero-pad to ``target_size``, no arithmetic -- and the float16 round trip of
data_loader.py:41-42 is DEFERRED to the consuming process: ``finish_on_device`` (called by this package's
``train.train_one_epoch`` / ``validate_one_epoch`` on every batch) applies it with the CUDA kernel.  The kernel is idempotent
on already-rounded values, so batches from ``num_workers=0`` (rounded in ``__getitem__``) pass through unchanged.  A script
that consumes worker batches without ``finish_on_device`` gets un-rounded (float32-exact) spectrograms; set
``strict_workers=True`` to get a RuntimeError in workers instead.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.utils.data
from torch.utils.data import Dataset

from . import _lib


def spec_f16_crop(spec: torch.Tensor, target_size=(256, 64)) -> torch.Tensor:
    """(N, F, T) float32 CUDA -> (N, f_out, t_out) float32 CUDA: the data_loader.py:41-72 transform on the device."""
    _lib.require_cuda()
    if not spec.is_cuda:
        raise _lib.AdnError("spec_f16_crop needs a CUDA tensor (no CPU fallback)")
    if spec.dim() == 2:
        spec = spec.unsqueeze(0)
    spec = spec.float().contiguous()
    n, f_in, t_in = spec.shape
    f_out, t_out = int(target_size[0]), int(target_size[1])
    out = torch.empty((n, f_out, t_out), dtype=torch.float32, device=spec.device)
    with torch.cuda.device(spec.device):
        st = _lib.load().adn_spec_f16_crop_f32(spec.data_ptr(), n, f_in, t_in, f_out, t_out, out.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "adn_spec_f16_crop_f32")
    return out


def finish_on_device(batch: torch.Tensor, device=None) -> torch.Tensor:
    """The deferred half of SpectrogramDataset's transform for batches that came out of DataLoader workers: (B, 1, F, T) float32
    host or CUDA tensor -> CUDA tensor rounded through float16 (data_loader.py:41-42) by ``adn_spec_f16_crop_f32`` (same
    size in and out).  Idempotent: already-rounded batches are returned bit-identical."""
    _lib.require_cuda()
    dev = torch.device(device) if device is not None else (batch.device if batch.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    b = batch.to(dev, non_blocking=True)
    shape = b.shape
    return spec_f16_crop(b.reshape(-1, shape[-2], shape[-1]), (shape[-2], shape[-1])).reshape(shape)


class SpectrogramDataset(Dataset):
    """data_loader.py:7-52."""

    def __init__(self, data_dir, target_size=(256, 64), device_output: bool = False, strict_workers: bool = False):
        self.pairs = []
        self.target_size = target_size
        self.device_output = bool(device_output)      # extension: keep items on the GPU instead of copying them back
        self.strict_workers = bool(strict_workers)    # extension: raise inside DataLoader workers instead of deferring
        clean_files = sorted(os.path.join(data_dir, f) for f in os.listdir(data_dir) if f.startswith("clean") and f.endswith(".npy"))
        noisy_files = sorted(os.path.join(data_dir, f) for f in os.listdir(data_dir) if f.startswith("noisy") and f.endswith(".npy"))
        print(f"Found {len(clean_files)} clean files and {len(noisy_files)} noisy files in {data_dir}")
        assert len(clean_files) == len(noisy_files), f"Mismatch in {data_dir}"
        self.pairs.extend(zip(noisy_files, clean_files))
        print(f"Total pairs loaded: {len(self.pairs)}")

    def __len__(self):
        return len(self.pairs)

    def _load(self, path):
        a = np.load(path)
        if a.ndim != 2:
            raise ValueError(f"{path}: expected a 2-D (freq, time) spectrogram")
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    def _layout_only(self, spec: torch.Tensor) -> torch.Tensor:
        """The crop / zero-pad of data_loader.py:54-72 as pure slicing (no arithmetic) -> (1, f_out, t_out) float32 host tensor;
        the float16 round trip is left to ``finish_on_device`` in the consuming process."""
        f_out, t_out = int(self.target_size[0]), int(self.target_size[1])
        out = torch.zeros((1, f_out, t_out), dtype=torch.float32)
        f, t = min(f_out, spec.shape[0]), min(t_out, spec.shape[1])
        out[0, :f, :t] = spec[:f, :t]
        return out

    def __getitem__(self, idx):
        noisy_path, clean_path = self.pairs[idx]
        if torch.utils.data.get_worker_info() is not None:        # a DataLoader worker process: no CUDA here (module docstring)
            if self.strict_workers:
                raise RuntimeError("SpectrogramDataset: the float16 transform runs on the GPU and DataLoader workers have no CUDA "
                                   "context; use num_workers=0, load_batch(), or strict_workers=False + finish_on_device()")
            return self._layout_only(self._load(noisy_path)), self._layout_only(self._load(clean_path))
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if dev is None:
            _lib.require_cuda()
        noisy = spec_f16_crop(self._load(noisy_path).to(dev), self.target_size)       # (1, 256, 64)
        clean = spec_f16_crop(self._load(clean_path).to(dev), self.target_size)
        if self.device_output:
            return noisy, clean
        return noisy.cpu(), clean.cpu()

    def load_batch(self, indices):
        """Batched fast path: (B, 1, f, t) float32 CUDA tensors for the given item indices (files of equal shape are
        stacked and transformed by one launch each for noisy / clean)."""
        _lib.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        outs = []
        for col in (0, 1):
            arrs = [self._load(self.pairs[i][col]) for i in indices]
            if len({tuple(a.shape) for a in arrs}) == 1:
                outs.append(spec_f16_crop(torch.stack(arrs).to(dev), self.target_size).unsqueeze(1))
            else:
                outs.append(torch.stack([spec_f16_crop(a.to(dev), self.target_size) for a in arrs]))
        return outs[0], outs[1]

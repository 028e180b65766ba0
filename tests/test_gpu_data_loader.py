"""SpectrogramDataset drop-in against a numpy restatement of data_loader.py:37-72 (f16 round trip + pad/crop) -- bit-exact."""
import numpy as np
import pytest
import torch

from audiodenoiser_b200.data_loader import SpectrogramDataset, spec_f16_crop

pytestmark = pytest.mark.gpu


def ref_transform(a, target=(256, 64)):
    """data_loader.py:41-72 restated with numpy."""
    a = a.astype(np.float16)
    th, tw = target
    h, w = a.shape
    if h < th:
        a = np.pad(a, ((0, th - h), (0, 0)), mode="constant")
    elif h > th:
        a = a[:th, :]
    if w < tw:
        a = np.pad(a, ((0, 0), (0, tw - w)), mode="constant")
    elif w > tw:
        a = a[:, :tw]
    return a.astype(np.float32)[None]


@pytest.mark.parametrize("shape", [(257, 122), (257, 188), (257, 30), (100, 64), (256, 64), (300, 10)])
def test_transform_bit_exact(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    a = (np.abs(rng.standard_normal(shape)) * 10 ** rng.uniform(-6, 2, shape)).astype(np.float32)
    a[0, 0] = 70000.0       # overflows float16 -> inf, as numpy's astype does
    a[1, 1] = 1e-9          # underflows to 0 / subnormal
    got = spec_f16_crop(torch.from_numpy(a).cuda()).cpu().numpy()
    assert np.array_equal(got, ref_transform(a), equal_nan=True)


def test_dataset_pairs_and_items(tmp_path):
    rng = np.random.default_rng(0)
    for k in range(3):
        np.save(tmp_path / f"noisy_white_chunk_{k}.npy", np.abs(rng.standard_normal((257, 122))).astype(np.float32))
        np.save(tmp_path / f"clean_white_chunk_{k}.npy", np.asfortranarray(np.abs(rng.standard_normal((257, 122))).astype(np.float32)))
    ds = SpectrogramDataset(str(tmp_path))
    assert len(ds) == 3
    noisy, clean = ds[1]
    assert noisy.shape == (1, 256, 64) and noisy.dtype == torch.float32 and not noisy.is_cuda
    assert np.array_equal(noisy.numpy(), ref_transform(np.load(tmp_path / "noisy_white_chunk_1.npy")))
    assert np.array_equal(clean.numpy(), ref_transform(np.load(tmp_path / "clean_white_chunk_1.npy")))
    nb, cb = ds.load_batch([0, 2])
    assert nb.shape == (2, 1, 256, 64) and nb.is_cuda
    assert torch.equal(nb[1].cpu(), ds[2][0])
    (tmp_path / "clean_extra.npy").write_bytes(b"")
    with pytest.raises(AssertionError):
        SpectrogramDataset(str(tmp_path))


def test_dataloader_with_worker_processes(tmp_path):
    """train.py:118-119 wraps the dataset in DataLoader(num_workers=4, pin_memory=True).  Worker processes have no CUDA context:
    they only crop / pad (no arithmetic) and the float16 round trip is applied by finish_on_device in the consuming process;
    the finished batch is bit-identical to the main-process items, and finishing is idempotent."""
    from torch.utils.data import DataLoader, random_split
    from audiodenoiser_b200.data_loader import finish_on_device
    from audiodenoiser_b200.train import _deferred_transform
    rng = np.random.default_rng(1)
    for k in range(6):
        np.save(tmp_path / f"noisy_white_chunk_{k}.npy", (np.abs(rng.standard_normal((257, 122))) * 3).astype(np.float32))
        np.save(tmp_path / f"clean_white_chunk_{k}.npy", (np.abs(rng.standard_normal((257, 40))) * 3).astype(np.float32))
    ds = SpectrogramDataset(str(tmp_path))
    torch.cuda.init()                                         # the parent holds a CUDA context, as train.py:122 does before iterating
    loader = DataLoader(ds, batch_size=3, shuffle=False, num_workers=2, pin_memory=True)
    assert _deferred_transform(loader) and not _deferred_transform(DataLoader(ds, batch_size=3))
    sub, _ = random_split(ds, [4, 2])
    assert _deferred_transform(DataLoader(sub, batch_size=2, num_workers=2))
    seen = 0
    for noisy, clean in loader:
        assert noisy.shape == (3, 1, 256, 64) and not noisy.is_cuda
        fn, fc = finish_on_device(noisy), finish_on_device(clean)
        assert fn.is_cuda and torch.equal(finish_on_device(fn), fn)
        for j in range(3):
            mn, mc = ds[seen + j]                             # main process: the CUDA transform inside __getitem__
            assert torch.equal(fn[j].cpu(), mn) and torch.equal(fc[j].cpu(), mc)
        seen += 3
    assert seen == 6
    strict = SpectrogramDataset(str(tmp_path), strict_workers=True)
    with pytest.raises(RuntimeError, match="num_workers=0"):
        next(iter(DataLoader(strict, batch_size=2, num_workers=1)))


@pytest.mark.parametrize("length,center", [(16000, False), (24000, True), (4000, False), (8100, True)])
def test_stft_second_output_is_the_loader_transform(length, center):
    """the STFT kernel's second output == SpectrogramDataset's transform (float16 round trip, crop / zero-pad
    to (256, 64), data_loader.py:41-72) of its own magnitudes, bit for bit -- with and without the full-magnitude output."""
    from audiodenoiser_b200 import spectral, synth
    x = torch.from_numpy(np.stack([synth.make_clip(i, "R")[:length] for i in range(3)])).to(torch.device("cuda", 0))
    mag = spectral.stft_mag_batched(x, center)
    crop, mag2 = spectral.stft_mag_train_batched(x, center, with_mag=True)
    only = spectral.stft_mag_train_batched(x, center)
    # the plain |STFT| entry may run another kernel variant than the fused one (ptxas contracts packed mul + add pairs differently
    # per kernel: last-bit differences), so the transform is checked against the fused kernel's OWN magnitudes
    assert torch.allclose(mag2, mag, rtol=0.0, atol=2e-6 * float(mag.abs().max()))
    ref = np.zeros((3, 1, 256, 64), np.float32)
    m = mag2.cpu().numpy().astype(np.float16).astype(np.float32)
    f, t = min(256, m.shape[1]), min(64, m.shape[2])
    ref[:, 0, :f, :t] = m[:, :f, :t]
    assert crop.shape == (3, 1, 256, 64)
    assert np.array_equal(crop.cpu().numpy(), ref) and np.array_equal(only.cpu().numpy(), ref)

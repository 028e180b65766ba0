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


@pytest.mark.parametrize("length,center", [(16000, False), (24000, True), (4000, False), (8100, True)])
def test_stft_second_output_is_the_loader_transform(length, center):
    """the STFT kernel's second output == SpectrogramDataset's transform (float16 round trip, crop / zero-pad
    to (256, 64), data_loader.py:41-72) of its own magnitudes, bit for bit -- with and without the full-magnitude output."""
    from audiodenoiser_b200 import spectral, synth
    x = torch.from_numpy(np.stack([synth.make_clip(i, "R")[:length] for i in range(3)])).to(torch.device("cuda", 0))
    mag = spectral.stft_mag_batched(x, center)
    crop, mag2 = spectral.stft_mag_train_batched(x, center, with_mag=True)
    only = spectral.stft_mag_train_batched(x, center)
    assert torch.equal(mag2, mag)
    ref = np.zeros((3, 1, 256, 64), np.float32)
    m = mag.cpu().numpy().astype(np.float16).astype(np.float32)
    f, t = min(256, m.shape[1]), min(64, m.shape[2])
    ref[:, 0, :f, :t] = m[:, :f, :t]
    assert crop.shape == (3, 1, 256, 64)
    assert np.array_equal(crop.cpu().numpy(), ref) and np.array_equal(only.cpu().numpy(), ref)

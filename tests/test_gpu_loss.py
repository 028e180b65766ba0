"""CombinedPerceptualLoss drop-in (CUDA) against the values the reference's own loss.py produced (tests/golden/loss_*.npz, made
by oracle/make_golden.py through torchaudio) and against the fp32 oracle on other shapes.  Tolerance: 2e-5 relative
(fp32 DFT-vs-FFT rounding), the same bound the oracle itself is pinned with."""
import os

import numpy as np
import pytest
import torch

from audiodenoiser_b200 import _lib
from audiodenoiser_b200.loss import CombinedPerceptualLoss, MelSpectrogramLoss, MultiScaleSTFTLoss, mel_filterbank
from oracle import loss_oracle

pytestmark = pytest.mark.gpu
RTOL = 2e-5


@pytest.mark.parametrize("name", ["train", "test"])
def test_matches_reference_fixture(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"loss_{name}.npz"))
    p = torch.from_numpy(z["pred"].astype(np.float32)).cuda(); t = torch.from_numpy(z["target"].astype(np.float32)).cuda()
    vals = [float(v) for v in CombinedPerceptualLoss()(p, t)]
    assert np.allclose(vals, z["values"], rtol=RTOL, atol=1e-7)


def test_seed0_regression(golden_dir):
    torch.manual_seed(0)
    p, t = torch.rand(4, 1, 256, 64), torch.rand(4, 1, 256, 64)
    vals = [float(v) for v in CombinedPerceptualLoss()(p.cuda(), t.cuda())]
    assert np.allclose(vals, np.load(os.path.join(golden_dir, "loss_seed0.npz"))["values"], rtol=RTOL)


@pytest.mark.parametrize("shape", [(1, 1, 257, 188), (3, 1, 256, 64), (2, 1, 257, 1034), (5, 1, 17, 32), (2, 1, 64, 100)])
def test_matches_oracle_on_other_shapes(shape):
    g = torch.Generator().manual_seed(shape[2] + shape[3])
    p = torch.randn(shape, generator=g).abs() * 3; t = torch.randn(shape, generator=g).abs() * 3
    ref = [float(v) for v in loss_oracle.combined_loss(p, t)]
    crit = CombinedPerceptualLoss()
    got = crit(p.cuda(), t.cuda())
    assert all(v.is_cuda and v.dim() == 0 for v in got)
    assert np.allclose([float(v) for v in got], ref, rtol=RTOL, atol=1e-7)
    assert abs(float(got[0]) - (0.4 * float(got[1]) + 0.4 * float(got[2]) + 0.2 * float(got[3]))) <= 1e-6 * max(1.0, abs(float(got[0])))
    assert np.isclose(float(MultiScaleSTFTLoss()(p.cuda(), t.cuda())), ref[1], rtol=RTOL)
    assert np.isclose(float(MelSpectrogramLoss()(p.cuda(), t.cuda())), ref[2], rtol=RTOL)


def test_deterministic_and_identity():
    p = torch.rand(6, 1, 257, 188, device="cuda")
    a = CombinedPerceptualLoss()(p, p * 0.5); b = CombinedPerceptualLoss()(p, p * 0.5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    z = CombinedPerceptualLoss()(p, p)
    assert all(float(v) == 0.0 for v in z)


def test_filterbank_and_validation():
    assert torch.equal(mel_filterbank(), loss_oracle.mel_filterbank().float())
    crit = CombinedPerceptualLoss()
    assert (crit.w_stft, crit.w_mel, crit.w_l1) == (0.4, 0.4, 0.2)
    with pytest.raises(ValueError):
        crit(torch.rand(1, 1, 8, 64), torch.rand(1, 1, 9, 64))                                     # shape mismatch (host tensors)
    with pytest.raises(ValueError):
        crit(torch.rand(1, 1, 8, 16, device="cuda"), torch.rand(1, 1, 8, 16, device="cuda"))      # T <= 31: reflect pad impossible


def test_host_tensors_in_host_scalars_out(golden_dir):
    """test.py:119-121 passes CPU tensors and calls .item() on the four results: host in -> the same kernels -> host out."""
    z = np.load(os.path.join(golden_dir, "loss_test.npz"))
    crit = CombinedPerceptualLoss()
    p, t = torch.from_numpy(z["pred"].astype(np.float32)), torch.from_numpy(z["target"].astype(np.float32))
    host = crit(p, t)
    devv = crit(p.cuda(), t.cuda())
    assert all(not v.is_cuda and v.dim() == 0 for v in host)
    for a, b in zip(host, devv):
        assert float(a) == float(b)
    # mixed placement (prediction on the GPU, target still on the host) follows the prediction
    mixed = crit(p.cuda(), t)
    assert mixed[0].is_cuda and float(mixed[0]) == float(devv[0])

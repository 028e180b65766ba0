"""The C-ABI library: loads, exports every symbol include/adn_b200.h declares, and its host-only logic
(frame counts, error strings, argument validation that returns before any CUDA work).  CPU only."""
import ctypes
import os
import re

import numpy as np
import pytest

from audiodenoiser_b200 import _lib
from oracle import stft_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "adn_b200.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(adn_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_what_python_binds():
    assert header_functions() == sorted(_lib.SIGNATURES)


def test_library_loads_and_exports_every_symbol():
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(raw, name), name
    assert lib.adn_version() >= 100


def test_every_entry_point_cites_the_reference():
    """Each compute entry point names the reference call site it replaces (file:line)."""
    text = open(HEADER).read()
    for cite in ("create_train_dataset.py:162-174", "create_test_dataset.py:35-41", "test.py:36-37,40,48", "model.py:11-16",
                 "model.py:38,43", "model.py:68,93", "data_loader.py:37-72", "train.py:65-72", "train.py:70", "loss.py:83-95",
                 "create_train_dataset.py:105-159"):
        assert cite in text, cite


@pytest.mark.parametrize("length", [0, 1, 127, 128, 511, 512, 513, 639, 640, 16000, 24000, 24001, 132300])
@pytest.mark.parametrize("center", [0, 1])
def test_num_frames_matches_oracle(length, center):
    lib = _lib.load()
    got = lib.adn_stft_num_frames(length, center)
    if not center and length < 512:
        assert got == -1
        with pytest.raises(ValueError):
            stft_oracle.num_frames(length, bool(center))
    else:
        assert got == stft_oracle.num_frames(length, bool(center))


def test_status_strings_and_host_side_argument_errors():
    lib = _lib.load()
    for st in range(0, 6):
        assert lib.adn_error_string(st)
    # too-short input with center=0 is rejected before any device work (librosa raises ParameterError there)
    x = np.zeros(100, np.float32)
    out = np.zeros(257, np.float32)
    st = lib.adn_stft_mag_host_f32(x.ctypes.data_as(ctypes.c_void_p), 1, 100, 0, out.ctypes.data_as(ctypes.c_void_p))
    assert st == 5
    with pytest.raises(ValueError):
        _lib.check(st, "adn_stft_mag_host_f32")
    assert lib.adn_stft_mag_host_f32(None, -1, 1000, 1, None) == 1
    assert lib.adn_istft_ola_host_f32(None, None, 0, 1, 0, None) == 1
    # empty batch is a no-op success
    assert lib.adn_stft_mag_host_f32(None, 0, 1000, 1, None) == 0


def test_product_path_raises_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from audiodenoiser_b200 import create_test_dataset, model, spectral
    with pytest.raises(_lib.AdnError):
        create_test_dataset.audio_to_spectrogram(np.zeros(4000, np.float32))
    with pytest.raises(_lib.AdnError):
        spectral.stft_mag_batched(torch.zeros(1, 4000))
    net = model.UNet().eval()
    with pytest.raises(_lib.AdnError):
        net(torch.zeros(1, 1, 32, 32))
    # the training step, the loss, the noise mixing and the fused loader transform have no CPU path either
    from audiodenoiser_b200 import create_train_dataset, loss, training
    with pytest.raises(_lib.AdnError):
        training.TrainEngine(model.UNet())
    with pytest.raises(_lib.AdnError):
        net.train()(torch.zeros(2, 1, 32, 32))
    with pytest.raises(_lib.AdnError):
        loss.CombinedPerceptualLoss()(torch.zeros(1, 1, 64, 64), torch.zeros(1, 1, 64, 64))
    with pytest.raises(_lib.AdnError):
        create_train_dataset.add_noise(np.zeros(16000, np.float32), None, "white")
    with pytest.raises(_lib.AdnError):
        spectral.stft_mag_train_batched(torch.zeros(1, 16000))


def test_training_entry_points_reject_bad_arguments_before_any_device_work():
    lib = _lib.load()
    assert lib.adn_train_workspace_bytes() > 0 and lib.adn_wgrad_workspace_bytes() > 0
    assert lib.adn_loss_backward_workspace_bytes(4, 256, 64) > 0 and lib.adn_loss_backward_workspace_bytes(-1, 256, 64) == -1
    assert lib.adn_bn_relu_apply_bf16(None, None, None, 10, 64, None, None) == 1
    assert lib.adn_maxpool2x2_backward_add_bf16(None, None, None, 0, 1, 4, 4, 64, None, None) == 1
    assert lib.adn_adamw_step_f32(None, None, None, None, 0, None, 1e-4, 0.9, 0.999, 1e-8, 0.01, 1, None) == 1
    assert lib.adn_grad_norm_f32(None, 0, 1.0, None, None, None) == 1
    assert lib.adn_mix_noise_snr_f32(None, None, 0, 16000, 8.0, None, None) == 0        # empty batch
    assert lib.adn_mix_noise_snr_f32(None, None, 2, 16000, 8.0, None, None) == 1
    assert lib.adn_pack_weights_table_bf16(None, 3, None) == 1


def test_flat_parameter_layout_is_the_state_dict_order():
    """TrainEngine's flat buffers: every trainable tensor of the 136-key state_dict, in order, each slice 256-byte aligned."""
    from audiodenoiser_b200.checkpoint import state_dict_spec
    from audiodenoiser_b200.training import flat_layout
    offsets, total = flat_layout()
    spec = state_dict_spec()
    trainable = [k for k in spec if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    assert list(offsets) == trainable and len(trainable) == 82
    end = 0
    n_params = 0
    for k, (off, numel, shape) in offsets.items():
        assert off % 64 == 0 and off >= end and tuple(shape) == tuple(spec[k][0])
        end = off + numel
        n_params += numel
    assert n_params == 31042369 and total >= end and total % 64 == 0

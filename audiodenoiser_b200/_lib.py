"""ctypes binding of libadn_b200.so (C ABI in include/adn_b200.h).

There is no CPU fallback: if the library cannot be loaded, or the current device is not a B200-class GPU,
every call raises.  The library is built in-tree by ``python -m audiodenoiser_b200.build``.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_float, c_int, c_int64, c_uint64, c_void_p, c_char_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libadn_b200.so")

_lock = threading.Lock()
_lib = None

P = c_void_p

# name -> (restype, argtypes); mirrors include/adn_b200.h one to one
SIGNATURES = {
    "adn_version": (c_int, []),
    "adn_source_hash": (c_char_p, []),
    "adn_error_string": (c_char_p, [c_int]),
    "adn_last_cuda_error": (c_int, []),
    "adn_device_check": (c_int, []),
    "adn_stft_num_frames": (c_int64, [c_int64, c_int]),
    "adn_stft_mag_f32": (c_int, [P, c_int64, c_int64, c_int64, c_int, P, P]),
    "adn_stft_mag_crop_f16_f32": (c_int, [P, c_int64, c_int64, c_int64, c_int, P, P, c_int, c_int, P]),
    "adn_stft_complex_f32": (c_int, [P, c_int64, c_int64, c_int64, c_int, P, P]),
    "adn_istft_ola_f32": (c_int, [P, P, c_int, c_uint64, c_int64, c_int64, P, P]),
    "adn_istft_ola_counter_f32": (c_int, [P, c_uint64, P, c_int64, c_int64, P, P]),
    "adn_u64_add": (c_int, [P, c_uint64, P]),
    "adn_random_phasor_c64": (c_int, [c_uint64, c_int64, c_int64, P, P]),
    "adn_stft_mag_host_f32": (c_int, [P, c_int64, c_int64, c_int, P]),
    "adn_istft_ola_host_f32": (c_int, [P, P, c_uint64, c_int64, c_int64, P]),
    "adn_mix_noise_snr_f32": (c_int, [P, P, c_int64, c_int64, c_float, P, P]),
    "adn_mix_noise_cancel_f32": (c_int, [P, P, c_int64, c_int64, c_int, c_int, c_float, P, P]),
    "adn_resample_poly_f32": (c_int, [P, c_int64, c_int, c_int64, c_int, c_int, P, c_int, c_int, c_int64, c_int64, P, P]),
    "adn_pack_conv3x3_weight_bf16": (c_int, [P, c_int, c_int, P, P]),
    "adn_pack_convt2x2_weight_bf16": (c_int, [P, c_int, c_int, P, P]),
    "adn_fold_bn_f32": (c_int, [P, P, P, P, P, c_float, c_int, P, P, P]),
    "adn_conv3x3_c1_bn_relu_bf16": (c_int, [P, c_int, c_int, c_int, P, P, P, P, P]),
    "adn_conv3x3_bn_relu_bf16": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P, P, P, P, P]),
    "adn_conv3x3_bn_relu_head_f32": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P, P, P, P, P, P]),
    "adn_convt2x2_bf16": (c_int, [P, c_int, c_int, c_int, c_int, P, c_int, P, P, P]),
    "adn_conv3x3_upmerged_bn_relu_bf16": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P, P, P, P, P]),
    "adn_pack_upmerged_weight_bf16": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P]),
    "adn_upmerged_pair_weight_elems": (c_int64, [c_int, c_int, c_int, c_int]),
    "adn_pack_upmerged_pair_weight_bf16": (c_int, [P, c_int, c_int, c_int, P, P, P]),
    "adn_conv3x3_upmerged_pair_bn_relu_bf16": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P, P, P]),
    "adn_conv3x3_pair_bn_relu_pool_bf16": (c_int, [P, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P, P, P]),
    "adn_maxpool2x2_bf16": (c_int, [P, c_int, c_int, c_int, c_int, P, P]),
    "adn_nhwc_bf16_to_nchw_f32": (c_int, [P, c_int, c_int, c_int, c_int, P, P]),
    "adn_nchw_f32_to_nhwc_bf16": (c_int, [P, c_int, c_int, c_int, c_int, P, P]),
    "adn_spec_f16_crop_f32": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, P, P]),
    "adn_spec_error_sums_f64": (c_int, [P, P, c_int64, P, P]),
    "adn_stats_pack_f64": (c_int, [P, c_int64, c_int64, P, P]),
    "adn_zero_bytes": (c_int, [P, c_int64, P]),
    "adn_copy_bytes": (c_int, [P, P, c_int64, P]),
    "adn_i64_add_n": (c_int, [P, c_int, c_int64, P]),
    "adn_loss_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "adn_combined_loss_f32": (c_int, [P, P, c_int64, c_int, c_int, P, P, P, P]),
    # training step (train.py:65-72)
    "adn_train_workspace_bytes": (c_int64, []),
    "adn_conv3x3_affine_bf16": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P, P, c_int, P, P]),
    "adn_conv3x3_c1_affine_bf16": (c_int, [P, c_int, c_int, c_int, P, P, P, c_int, P, P]),
    "adn_pack_conv3x3_dgrad_weight_bf16": (c_int, [P, c_int, c_int, P, P]),
    "adn_pack_convt2x2_dgrad_weight_bf16": (c_int, [P, c_int, c_int, P, P]),
    "adn_pack_weights_table_bf16": (c_int, [P, c_int, P]),
    "adn_pack_weights_table_sel_bf16": (c_int, [P, c_int, c_int, P]),
    "adn_pack_weights_table_flat_bf16": (c_int, [P, c_int, c_int, c_int, P]),
    "adn_bn_train_stats_f32": (c_int, [P, c_int64, c_int, P, P, c_float, c_float, P, P, P, P, P, P, P, P]),
    "adn_bn_relu_apply_bf16": (c_int, [P, P, P, c_int64, c_int, P, P]),
    "adn_bn_relu_backward_bf16": (c_int, [P, c_int, P, c_int64, c_int, P, P, P, P, P, P, P, P, P]),
    "adn_channel_sum_f32": (c_int, [P, c_int, c_int64, c_int, P, P, P]),
    "adn_maxpool2x2_backward_add_bf16": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P]),
    "adn_head1x1_forward_f32": (c_int, [P, P, P, c_int64, P, P]),
    "adn_head1x1_backward": (c_int, [P, P, P, c_int64, P, P, P, P, P]),
    "adn_wgrad_workspace_bytes": (c_int64, []),
    "adn_conv3x3_wgrad_f32": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, c_int, P, P]),
    "adn_conv3x3_c1_wgrad_f32": (c_int, [P, P, c_int, c_int, c_int, P, P, P]),
    "adn_convt2x2_wgrad_f32": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "adn_convt2x2_dgrad_bf16": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P, P]),
    "adn_loss_backward_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "adn_combined_loss_backward_f32": (c_int, [P, P, c_int64, c_int, c_int, P, c_float, c_float, c_float, P, P, P]),
    "adn_grad_norm_f32": (c_int, [P, c_int64, c_float, P, P, P]),
    "adn_adamw_step_f32": (c_int, [P, P, P, P, c_int64, P, c_float, c_float, c_float, c_float, c_float, c_int64, P]),
    "adn_adamw_step_dev_f32": (c_int, [P, P, P, P, c_int64, P, P, c_float, c_float, c_float, c_float, c_float, P]),
}


class AdnError(RuntimeError):
    pass


def _built_hash() -> str | None:
    """The source hash stamped into the .so on disk, read WITHOUT loading it into this process (a loaded library cannot be
    replaced): a throw-away python reads ``adn_source_hash()``."""
    import subprocess
    import sys
    code = ("import ctypes,sys;l=ctypes.CDLL(sys.argv[1]);f=l.adn_source_hash;f.restype=ctypes.c_char_p;print(f().decode())")
    try:
        r = subprocess.run([sys.executable, "-c", code, LIB_PATH], capture_output=True, text=True, timeout=120)
    except Exception:  # noqa: BLE001
        return None
    return r.stdout.strip() if r.returncode == 0 and r.stdout.strip() else None


def load(build_if_missing: bool = True):
    """Load (once) and return the ctypes handle.  Raises AdnError when the native library is unavailable, or when it was
    built from other sources than the ones beside it and cannot be rebuilt (no silent use of a stale binary)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        want = _build.source_hash()
        stamp = os.path.join(_build.BUILD, "source_hash.txt")
        stamped = open(stamp).read().strip() if os.path.exists(stamp) else None
        fresh = os.path.exists(LIB_PATH) and stamped == want and os.path.getmtime(LIB_PATH) >= os.path.getmtime(stamp) - 1.0
        if not fresh and os.path.exists(LIB_PATH) and _built_hash() == want:
            fresh = True
        if not fresh and build_if_missing:
            try:
                _build.build()                       # mtime- and hash-incremental; a no-op when everything is current
            except Exception as exc:  # noqa: BLE001
                raise AdnError(f"libadn_b200.so is missing or stale and could not be built: {exc}") from exc
        if not os.path.exists(LIB_PATH):
            raise AdnError(f"{LIB_PATH} not found: run `python -m audiodenoiser_b200.build` (no CPU fallback exists)")
        try:
            import torch  # noqa: F401  (brings libcudart.so.12 into the process: the library links the shared runtime)
        except Exception:  # noqa: BLE001
            pass
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as exc:
            raise AdnError(f"cannot load {LIB_PATH}: {exc}") from exc
        try:
            lib.adn_source_hash.restype = c_char_p
            have = lib.adn_source_hash().decode()
        except AttributeError:
            have = "missing"
        if have != want:
            raise AdnError(f"{LIB_PATH} was built from other sources (hash {have}, sources {want}): run `python -m audiodenoiser_b200.build`")
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise AdnError(f"{LIB_PATH} does not export {name}; rebuild it") from exc
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    """Translate an adn_status into a Python exception (ValueError for argument errors, AdnError otherwise)."""
    if status == 0:
        return
    lib = load()
    msg = lib.adn_error_string(status).decode()
    if status == 2:
        msg += f" [cudaError {lib.adn_last_cuda_error()}]"
    text = f"{what}: {msg}" if what else msg
    if status in (1, 5):
        raise ValueError(text)
    raise AdnError(text)


def require_cuda():
    """The product path needs a CUDA device; say so plainly instead of falling back."""
    import torch
    if not torch.cuda.is_available():
        raise AdnError("audiodenoiser_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def stream_ptr(torch_stream=None) -> int:
    import torch
    s = torch_stream if torch_stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)

"""Batched device-resident spectral front / back end (SURVEY 8b "batched fast path").

``stft_mag_batched`` / ``istft_batched`` take and return CUDA tensors and launch the sm_100a kernels of
libadn_b200.so on the current torch stream.  torch is used only for device memory and streams.
"""
from __future__ import annotations

from . import _lib

N_FFT = 512
HOP_LENGTH = 128
N_BINS = N_FFT // 2 + 1


def num_frames(length: int, center: bool) -> int:
    """1 + (L [+ 512] - 512)//128 -- librosa.stft's frame count; raises like librosa when too short."""
    t = _lib.load().adn_stft_num_frames(int(length), int(bool(center)))
    if t < 0:
        raise ValueError(f"Input signal length={length} is too small to analyze with n_fft={N_FFT}")
    return int(t)


def _check_wave(wave):
    torch = _lib.require_cuda()
    if not isinstance(wave, torch.Tensor) or not wave.is_cuda:
        raise _lib.AdnError("expected a CUDA tensor (no CPU fallback)")
    if wave.dim() == 1:
        wave = wave.unsqueeze(0)
    if wave.dim() != 2:
        raise ValueError("wave must be (L,) or (N, L)")
    if wave.dtype != torch.float32:
        wave = wave.float()
    if wave.stride(1) != 1 or (wave.shape[0] > 1 and wave.stride(0) < wave.shape[1]):
        wave = wave.contiguous()
    return torch, wave


def stft_mag_batched(wave, center: bool = True, out=None):
    """(N, L) float32 CUDA -> (N, 257, T) float32 |STFT| (n_fft=512, hop=128, periodic Hann).
    center=True zero-pads 256 samples each side (create_test_dataset.py:39); center=False is the
    train-path framing (create_train_dataset.py:167-172)."""
    torch, wave = _check_wave(wave)
    n, length = wave.shape
    t = num_frames(length, center)
    if out is None:
        out = torch.empty((n, N_BINS, t), dtype=torch.float32, device=wave.device)
    elif tuple(out.shape) != (n, N_BINS, t) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 (N,257,T) tensor")
    stride = wave.stride(0) if n > 1 else max(length, 1)
    with torch.cuda.device(wave.device):
        st = _lib.load().adn_stft_mag_f32(wave.data_ptr(), n, length, stride, int(bool(center)), out.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "adn_stft_mag_f32")
    return out


def stft_mag_train_batched(wave, center: bool = False, target_size=(256, 64), with_mag: bool = False):
    """(N, L) float32 CUDA -> the training tensor of SpectrogramDataset.__getitem__ (data_loader.py:37-72), (N, 1, 256, 64)
    float32: |STFT| rounded through float16 and zero-padded / cropped to ``target_size``, produced by the STFT kernel itself.
    ``with_mag=True`` also returns the full (N, 257, T) magnitudes (the .npy the reference saves, create_train_dataset.py:251-254);
    without it only the frames that survive the crop are transformed."""
    torch, wave = _check_wave(wave)
    n, length = wave.shape
    t = num_frames(length, center)
    f_out, t_out = int(target_size[0]), int(target_size[1])
    crop = torch.empty((n, 1, f_out, t_out), dtype=torch.float32, device=wave.device)
    mag = torch.empty((n, N_BINS, t), dtype=torch.float32, device=wave.device) if with_mag else None
    stride = wave.stride(0) if n > 1 else max(length, 1)
    with torch.cuda.device(wave.device):
        st = _lib.load().adn_stft_mag_crop_f16_f32(wave.data_ptr(), n, length, stride, int(bool(center)), mag.data_ptr() if with_mag else 0,
                                                   crop.data_ptr(), f_out, t_out, _lib.stream_ptr())
    _lib.check(st, "adn_stft_mag_crop_f16_f32")
    return (crop, mag) if with_mag else crop


def stft_complex_batched(wave, center: bool = True):
    """(N, L) float32 CUDA -> (N, 257, T) complex64 STFT (librosa.stft at test.py:41)."""
    torch, wave = _check_wave(wave)
    n, length = wave.shape
    t = num_frames(length, center)
    out = torch.empty((n, N_BINS, t), dtype=torch.complex64, device=wave.device)
    stride = wave.stride(0) if n > 1 else max(length, 1)
    with torch.cuda.device(wave.device):
        st = _lib.load().adn_stft_complex_f32(wave.data_ptr(), n, length, stride, int(bool(center)), out.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "adn_stft_complex_f32")
    return out


def istft_batched(mag, phasor=None, seed: int = 0, out=None, phase_from=None, seed_counter=None):
    """Inverse STFT + overlap-add: (N,257,T) magnitude x (N,257,T) complex64 unit phasor -> (N, 128*(T-1)) float32.

    ``phasor=None`` draws a uniform random phase on the device from ``seed`` (test.py:36 uses the unseeded numpy RNG).
    A complex ``mag`` is taken as the complex spectrogram itself (librosa.istft at test.py:40).
    ``phase_from`` (opt-in, NOT the reference's behaviour): a complex (N,257,T) spectrogram whose phase is used, i.e. the
    kernel inverts ``mag * phase_from / |phase_from|`` (denoised magnitude + noisy phase).
    ``seed_counter``: a 1-element int64 CUDA tensor added to ``seed`` on the device (random-phase mode only), so that replays of a
    captured CUDA graph still draw a fresh phase per call; advance it with ``advance_counter``."""
    torch = _lib.require_cuda()
    if not isinstance(mag, torch.Tensor) or not mag.is_cuda:
        raise _lib.AdnError("expected a CUDA tensor (no CPU fallback)")
    if mag.dim() == 2:
        mag = mag.unsqueeze(0)
    if mag.dim() != 3 or mag.shape[1] != N_BINS:
        raise ValueError("spectrogram must be (257, T) or (N, 257, T)")
    n, _, t = mag.shape
    is_complex = mag.is_complex()
    mode = 1 if is_complex else 0
    if phase_from is not None:
        if is_complex or phasor is not None:
            raise ValueError("phase_from needs a real magnitude and no phasor")
        if phase_from.dim() == 2:
            phase_from = phase_from.unsqueeze(0)
        if tuple(phase_from.shape) != tuple(mag.shape) or not phase_from.is_complex():
            raise ValueError("phase_from must be a complex tensor of the magnitude's shape")
        mag = mag.float().contiguous()
        phase_from = phase_from.to(device=mag.device, dtype=torch.complex64).contiguous()
        mag_ptr, ph_ptr, mode = mag.data_ptr(), phase_from.data_ptr(), 2
    elif is_complex:
        if phasor is not None:
            raise ValueError("phasor must be None when the spectrogram is complex")
        spec = mag.to(torch.complex64).contiguous()
        mag_ptr, ph_ptr = 0, spec.data_ptr()
    else:
        mag = mag.float().contiguous()
        mag_ptr = mag.data_ptr()
        if phasor is not None:
            if phasor.dim() == 2:
                phasor = phasor.unsqueeze(0)
            if tuple(phasor.shape) != tuple(mag.shape):
                raise ValueError("phasor must have the magnitude's shape")
            phasor = phasor.to(device=mag.device, dtype=torch.complex64).contiguous()
            ph_ptr = phasor.data_ptr()
        else:
            ph_ptr = 0
    n_out = HOP_LENGTH * (t - 1)
    if out is None:
        out = torch.empty((n, n_out), dtype=torch.float32, device=mag.device)
    elif tuple(out.shape) != (n, n_out) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 (N, 128*(T-1)) tensor")
    with torch.cuda.device(mag.device):
        if seed_counter is not None and mode == 0 and ph_ptr == 0:
            if seed_counter.dtype != torch.int64 or not seed_counter.is_cuda or seed_counter.numel() != 1:
                raise ValueError("seed_counter must be a 1-element int64 CUDA tensor")
            st = _lib.load().adn_istft_ola_counter_f32(mag_ptr, int(seed) & 0xFFFFFFFFFFFFFFFF, seed_counter.data_ptr(), n, t,
                                                       out.data_ptr(), _lib.stream_ptr())
        else:
            st = _lib.load().adn_istft_ola_f32(mag_ptr, ph_ptr, mode, int(seed) & 0xFFFFFFFFFFFFFFFF, n, t,
                                               out.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "adn_istft_ola_f32")
    return out


def advance_counter(counter, inc: int = 1):
    """counter += inc on the device with a one-thread kernel of the library (capturable into a CUDA graph)."""
    torch = _lib.require_cuda()
    with torch.cuda.device(counter.device):
        _lib.check(_lib.load().adn_u64_add(counter.data_ptr(), int(inc), _lib.stream_ptr()), "adn_u64_add")
    return counter


def random_phasor(seed: int, n_clips: int, n_frames: int, device=None):
    """The (n_clips, 257, T) complex64 unit phasor ``istft_batched(mag, None, seed)`` applies: the seeded device-side
    stand-in for ``np.exp(2j * np.pi * np.random.rand(*mag.shape))`` (test.py:36)."""
    torch = _lib.require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((int(n_clips), N_BINS, int(n_frames)), dtype=torch.complex64, device=device)
    with torch.cuda.device(device):
        st = _lib.load().adn_random_phasor_c64(int(seed) & 0xFFFFFFFFFFFFFFFF, int(n_clips), int(n_frames), out.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "adn_random_phasor_c64")
    return out

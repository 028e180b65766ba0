"""End-to-end hot path on one GPU: noisy waveform -> |STFT| -> UNet -> inverse STFT / overlap-add.

This is the batched composition of the three drop-in pieces (create_test_dataset.audio_to_spectrogram, model.UNet.forward,
test.griffin_lim_reconstruction) with every intermediate resident in HBM and, for a fixed batch shape, the whole launch
sequence captured in one CUDA graph.  Reconstruction uses a random unit phasor (seeded, generated on the device) exactly as
the reference does (test.py:36), or the phasor the caller injects.
"""
from __future__ import annotations

import torch

from . import _lib, resample, spectral
from .model import UNet


class Denoiser:
    def __init__(self, model: UNet, center: bool = True, seed: int = 0, use_graph: bool = True, phase: str = "random",
                 input_sr: int | None = None):
        """phase="random" is the reference's reconstruction (test.py:36: a fresh random phase).  phase="noisy" (opt-in, SURVEY 8f
        row 4) reuses the phase of the noisy input's STFT -- audibly better, but NOT what the reference computes.
        input_sr: native sample rate of the input clips; when it is not 8000 the clips -- (N, L) mono or (N, C, L) planar -- are
        mixed down and resampled to 8 kHz on the device first (the tail of librosa.load(sr=8000), test.py:80)."""
        _lib.require_cuda()
        if phase not in ("random", "noisy"):
            raise ValueError("phase must be 'random' or 'noisy'")
        self.phase = phase
        self.input_sr = int(input_sr) if input_sr else None
        self.model = model.eval()
        self.center = bool(center)
        self.seed = int(seed)
        self.use_graph = bool(use_graph)
        self._graphs = {}
        self.launches_per_call = None
        # Random-phase reconstruction draws a FRESH phase per call, as test.py:36 does: call k uses seed + k.  The per-call part
        # lives in a device counter (advanced by a one-thread kernel at the end of every call, inside the captured graph too),
        # because a graph replay bakes kernel arguments in.  `last_seed` is the effective seed of the most recent call:
        # spectral.random_phasor(last_seed, ...) reproduces its phasor.
        self.calls = 0
        self.last_seed = self.seed
        self._counter = None

    def denoise(self, wave: torch.Tensor, phasor: torch.Tensor | None = None, return_spectrograms: bool = False):
        """wave (N, L) float32 -> audio (N, 128*(T-1)) float32 CUDA.  ``wave`` is a CUDA tensor, or a (pinned) host tensor
        that is copied straight into the graph's input buffer.  With return_spectrograms also returns
        (noisy_mag, denoised_mag), each (N, 257, T).  In graph mode the returned tensors are the graph's static output
        buffers: they are overwritten by the next call with the same shape."""
        if wave.dim() == 1:
            wave = wave.unsqueeze(0)
        if wave.dim() == 3 and not self.input_sr:
            raise ValueError("multi-channel input needs input_sr (the mono mix is part of the resample front-end)")
        if wave.dtype != torch.float32:
            wave = wave.float()
        if not self.use_graph:
            if not wave.is_cuda:
                wave = wave.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
            wave = wave.contiguous()
            audio, mag, den = self._run(wave, phasor)
            self._count_call(phasor)
            return (audio, mag, den) if return_spectrograms else audio
        dev = wave.device if wave.is_cuda else torch.device("cuda", torch.cuda.current_device())
        key = (tuple(wave.shape), str(dev), phasor is not None, self.phase, self.model._version_key(dev))
        g = self._graphs.get(key)
        if g is None:
            g = self._capture(wave.to(dev), phasor)
            self._graphs = {key: g}            # one resident shape
        with torch.cuda.device(dev):
            g["wave"].copy_(wave, non_blocking=True)
            if phasor is not None:
                g["phasor"].copy_(phasor.to(torch.complex64), non_blocking=True)
            g["graph"].replay()
        self._count_call(phasor)
        if return_spectrograms:
            return g["audio"], g["mag"], g["den"]
        return g["audio"]

    def _count_call(self, phasor):
        if phasor is None and self.phase == "random":
            self.last_seed = self.seed + self.calls
            self.calls += 1

    def _run(self, wave, phasor):
        if self.input_sr and (self.input_sr != resample.SAMPLE_RATE or wave.dim() == 3):
            wave = resample.resample_batched(wave, self.input_sr, resample.SAMPLE_RATE)
        if self.phase == "noisy" and phasor is None:
            spec = spectral.stft_complex_batched(wave, self.center)
            mag = spec.abs()
            den = self.model(mag.unsqueeze(1)).squeeze(1)
            return spectral.istft_batched(den, phase_from=spec), mag, den
        nvtx = torch.cuda.nvtx                               # ranges for nsys / ncu timelines; no-ops otherwise
        nvtx.range_push("adn.stft_mag")
        mag = spectral.stft_mag_batched(wave, self.center)
        nvtx.range_pop()
        nvtx.range_push("adn.unet_forward")
        den = self.model(mag.unsqueeze(1)).squeeze(1)
        nvtx.range_pop()
        nvtx.range_push("adn.istft_ola")
        try:
            return self._reconstruct(den, phasor, wave.device), mag, den
        finally:
            nvtx.range_pop()

    def _reconstruct(self, den, phasor, device):
        if phasor is not None:
            return spectral.istft_batched(den, phasor)
        if self._counter is None or self._counter.device != device:
            self._counter = torch.full((1,), self.calls, dtype=torch.int64, device=device)
        audio = spectral.istft_batched(den, None, seed=self.seed, seed_counter=self._counter)
        spectral.advance_counter(self._counter)
        return audio

    def _capture(self, wave, phasor):
        dev = wave.device
        static_wave = wave.contiguous().clone()
        static_ph = phasor.to(torch.complex64).clone() if phasor is not None else None
        with torch.cuda.device(dev):
            # warm-up on a side stream: lazy init (weight packing, workspace, func attributes) must not be captured
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    audio, mag, den = self._run(static_wave, static_ph)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            if self._counter is not None:
                self._counter.fill_(self.calls)            # the warm-up calls advanced the device counter: they do not count
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                audio, mag, den = self._run(static_wave, static_ph)
        # the graph holds raw pointers into the model's packed weights and activation workspace: keep them alive here even if
        # the model later repacks (new checkpoint) or switches to another input shape
        keep = (self.model._packed, dict(self.model._ws))
        return {"graph": graph, "wave": static_wave, "phasor": static_ph, "mag": mag, "den": den, "audio": audio, "keep": keep}

    # ------------------------------------------------------------------ host-buffer entry (what a script calls)
    def denoise_host(self, wave_host: torch.Tensor, out_host: torch.Tensor | None = None, device=None) -> torch.Tensor:
        """(N, L) float32 host tensor (pinned for full speed) -> (N, 128*(T-1)) float32 host tensor.  Host->device copy of
        the input and device->host copy of the result are part of the call; returns after the stream has drained."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(dev):
            audio = self.denoise(wave_host)
            if out_host is None:
                out_host = torch.empty(audio.shape, dtype=torch.float32, pin_memory=True)
            out_host.copy_(audio, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return out_host


def stream_host_batches(step_fn, host_inputs, host_outputs, host_stats=None, device=None):
    """Run ``step_fn(i, wave_dev) -> (audio_dev, stats_dev | None)`` over a sequence of HOST batches with the host<->device
    copies overlapped with compute: batch i+1 is copied in and batch i-1 copied out on a second stream while batch i runs.

    host_inputs / host_outputs: lists of pinned host tensors (one per batch; outputs are filled in place); host_stats: optional
    list of pinned tensors receiving each step's small statistics vector.  Every batch's host->device and device->host copy
    still happens inside this call -- it is the script-facing entry for a stream of batches, not a way to skip the copies.
    Two device staging buffers per direction; returns after both streams have drained."""
    _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = len(host_inputs)
    if n == 0:
        return host_outputs
    with torch.cuda.device(dev):
        comp = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        stage_in = [torch.empty(host_inputs[0].shape, dtype=host_inputs[0].dtype, device=dev) for _ in range(2)]
        stage_out, stage_stats = [None, None], [None, None]
        h2d = [torch.cuda.Event() for _ in range(n)]
        consumed = [None, None]          # compute finished reading stage_in[b]
        drained = [None, None]           # copy stream finished reading stage_out[b]
        copy.wait_stream(comp)
        with torch.cuda.stream(copy):
            stage_in[0].copy_(host_inputs[0], non_blocking=True)
            h2d[0].record(copy)
        for i in range(n):
            b = i & 1
            if i + 1 < n:
                with torch.cuda.stream(copy):
                    if consumed[b ^ 1] is not None:
                        copy.wait_event(consumed[b ^ 1])
                    stage_in[b ^ 1].copy_(host_inputs[i + 1], non_blocking=True)
                    h2d[i + 1].record(copy)
            comp.wait_event(h2d[i])
            audio, stats = step_fn(i, stage_in[b])
            consumed[b] = torch.cuda.Event(); consumed[b].record(comp)
            if stage_out[b] is None:
                stage_out[b] = torch.empty(host_outputs[i].shape, dtype=audio.dtype, device=dev)
                if stats is not None:
                    stage_stats[b] = torch.empty_like(stats)
            if drained[b] is not None:
                comp.wait_event(drained[b])
            stage_out[b].copy_(audio[: stage_out[b].shape[0]] if audio.shape != stage_out[b].shape else audio)
            if stats is not None:
                stage_stats[b].copy_(stats)
            done = torch.cuda.Event(); done.record(comp)
            with torch.cuda.stream(copy):
                copy.wait_event(done)
                host_outputs[i].copy_(stage_out[b], non_blocking=True)
                if stats is not None and host_stats is not None:
                    host_stats[i].copy_(stage_stats[b], non_blocking=True)
                drained[b] = torch.cuda.Event(); drained[b].record(copy)
        copy.synchronize()
        comp.synchronize()
    return host_outputs

#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json): audio-seconds denoised per second through
STFT -> UNet -> iSTFT on N B200s, with the kernel roofline and the reference's CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--variant B|R] [--batch 64]

One "step" = one pass of the hot path over one batch of synthetic clips per GPU (BASELINE config 3: batch 64, end-to-end
STFT -> net -> iSTFT; clips are 3 s @ 44.1 kHz = 132 300 samples -> (257,1034) spectrograms, variant "B", or the
reference-faithful 3 s @ 8 kHz = 24 000 samples -> (257,188), variant "R").  For N > 1 (torchrun, one rank per GPU)
every rank denoises its own shard (weak scaling) and the step ends with the NCCL all-gather of the denoised audio and
the all-reduce of the error statistics (BASELINE config 4).

JSON keys: see the driver contract.  `value` = device-timed whole-job throughput with inputs resident in HBM;
`e2e` = the same through Denoiser.denoise_host with pinned HOST buffers (H2D + D2H inside the timed region);
`roofline` = the tcgen05 implicit-GEMM conv kernel (the dominant kernel) against the measured bf16 peak;
`kernels` = per-kernel achieved figures incl. the STFT / iSTFT HBM GB/s; `cpu_baseline` = the CPU oracle chain
(the arithmetic the reference executes) on a bounded sample on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "audio-sec/sec denoised (STFT->net->iSTFT)"
UNIT = "audio-s/s"
CLIP_SECONDS = 3.0
GL_ITERATIONS = 50          # test.py:29 default


def unet_flops(h: int, w: int) -> float:
    """FLOPs of one UNet forward on a (1,1,h,w) input (conv 2*Ci*Co*9*H*W, convT 2*Ci*Co*4*Hin*Win, head 2*64*H*W)."""
    hs, ws = [h], [w]
    for _ in range(4):
        hs.append(hs[-1] // 2); ws.append(ws[-1] // 2)
    ch = [64, 128, 256, 512, 1024]
    f = 0.0
    cin = 1
    for l in range(5):
        f += 2.0 * 9 * hs[l] * ws[l] * (cin * ch[l] + ch[l] * ch[l])
        cin = ch[l]
    for l in (3, 2, 1, 0):
        f += 2.0 * 4 * hs[l + 1] * ws[l + 1] * ch[l + 1] * ch[l]                    # ConvTranspose2d
        f += 2.0 * 9 * hs[l] * ws[l] * (2 * ch[l] * ch[l] + ch[l] * ch[l])          # cat -> conv, conv
    f += 2.0 * 64 * h * w
    return f


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_burst": float(d["bf16_tflops"]),
                    "bf16_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
        except Exception:  # noqa: BLE001
            pass
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 100 ms while a timed region runs (NVML)."""

    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80}

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------- CPU reference chain
def cpu_reference_chain(waves, sd, threads: int):
    """The reference's CPU path for these clips, as the oracle restates it: librosa-style float64 STFT magnitude
    (create_test_dataset.py:39-40) -> model.py fp32 forward on torch CPU (test.py:112-113) -> griffin_lim_reconstruction
    with its literal 50 x (istft, stft) loop (test.py:29-48).  Returns seconds."""
    import numpy as np
    import torch
    from oracle import stft_oracle, unet_oracle
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    mags = np.stack([stft_oracle.stft_mag(w, True) for w in waves]).astype(np.float32)
    with torch.no_grad():
        den = unet_oracle.unet_forward(sd, torch.from_numpy(mags).unsqueeze(1)).squeeze(1).numpy()
    for d in den:
        stft_oracle.griffin_lim_reconstruction(d, 512, 128, iterations=GL_ITERATIONS)
    return time.perf_counter() - t0


TOL = {"stft_mag_max_rel": 1e-4, "unet_norm_rel": 1e-2, "audio_norm_rel": 1e-2, "snr_delta_db": 0.05}      # BASELINE.json north_star


def parity_on_benchmark_clip(wave, clean_wave, sd, seed, gpu_mag, gpu_den, gpu_audio, phasor):
    """In-run parity on the benchmark configuration: clip 0 of the timed batch through the CPU oracle
    (float64 STFT -> fp32 reference UNet -> iSTFT of den * phasor with the SAME phasor the device drew, exported by
    adn_random_phasor_c64) against what the GPU produced for that clip.  Returns the `parity` object of the JSON line."""
    import numpy as np
    import torch
    from oracle import stft_oracle, unet_oracle
    ref_mag = stft_oracle.stft_mag(wave.astype(np.float64), True).astype(np.float32)
    with torch.no_grad():
        ref_den = unet_oracle.unet_forward(sd, torch.from_numpy(ref_mag)[None, None]).numpy()[0, 0]
    ref_audio = stft_oracle.istft(ref_den.astype(np.float64) * phasor.astype(np.complex128))
    clean_mag = stft_oracle.stft_mag(clean_wave.astype(np.float64), True)
    nrel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))  # noqa: E731
    snr = lambda d: float(10.0 * np.log10(np.sum(clean_mag ** 2) / np.sum((clean_mag - d) ** 2)))  # noqa: E731
    out = {"clip": "clip 0 of timed batch 0 (same synthetic clip, same checkpoint, same injected phasor on both sides)",
           "stft_mag_max_rel": float(np.max(np.abs(gpu_mag - ref_mag)) / np.max(np.abs(ref_mag))),
           "frames": {"gpu": int(gpu_mag.shape[1]), "oracle": int(ref_mag.shape[1])},
           "unet_norm_rel": nrel(gpu_den, ref_den), "audio_norm_rel": nrel(gpu_audio, ref_audio),
           "snr_db": {"gpu": snr(gpu_den), "oracle": snr(ref_den)}, "tolerances": TOL, "oracle": "oracle/stft_oracle.py + oracle/unet_oracle.py (fp32 reference graph)"}
    out["snr_delta_db"] = abs(out["snr_db"]["gpu"] - out["snr_db"]["oracle"])
    out["pass"] = bool(out["frames"]["gpu"] == out["frames"]["oracle"] and all(out[k] <= TOL[k] for k in TOL))
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores.  The reference is
    pure Python over librosa / torch; librosa is absent from this image, so the arm runs the oracle port (numpy/scipy
    restatement of the librosa arithmetic + the torch-CPU UNet), with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import torch
    from audiodenoiser_b200 import synth
    from audiodenoiser_b200.checkpoint import seeded_state_dict
    cores = os.cpu_count() or 1
    sd = seeded_state_dict(7)
    per_step = max(1, args.ref_clips)
    n = per_step * (args.steps + args.warmup)
    clips = [synth.make_clip(i, args.variant) for i in range(min(n, 4))]
    steps = []
    for s in range(args.warmup + args.steps):
        waves = [clips[(s * per_step + j) % len(clips)] for j in range(per_step)]
        dt = cpu_reference_chain(waves, sd, cores)
        if s >= args.warmup:
            steps.append(dt)
    total = float(np.sum(steps))
    value = per_step * args.steps * CLIP_SECONDS / total
    t_frames = 1 + synth.VARIANTS[args.variant][1] // 128
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 STFT/iSTFT + f32 UNet (CPU)", "data": "synthetic",
        # the b200 arm's workload, as the contract asks; what ONE STEP of this arm really processes is stated beside it
        "config": {**workload_config(args, t_frames, args.batch),
                   "l2": "n/a (CPU arm)", "collectives": "none (one CPU process)",
                   "clips_timed_per_step": per_step,
                   "note": f"reference arm: ONE CPU process on this box's {cores} host cores; every step times {per_step} clip(s), a bounded sample of the "
                           f"{args.batch}-clip batch the b200 arm processes per GPU and step (value = audio-seconds of the timed clips / their time); at "
                           f"--gpus N > 1 it is still this one host process -- the reference has no multi-device path"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} clip(s)/step x {args.steps} steps, variant {args.variant}: float64 STFT -> fp32 torch-CPU UNet "
                                   f"({torch.get_num_threads()} threads) -> {GL_ITERATIONS}-iteration istft/stft loop + istft"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(args, t_frames, batch):
    sr, length = {"B": (44100, 132300), "R": (8000, 24000)}[args.variant]
    return {"workload": f"BASELINE config 3/4: end-to-end STFT->UNet->iSTFT, batch {batch} clips per GPU per step, "
                        f"3 s @ {sr} Hz synthetic tone/music + white/pink noise @ 8 dB",
            "variant": args.variant, "clip_samples": length, "spectrogram": [257, t_frames], "batch_per_gpu": batch,
            "n_fft": 512, "hop": 128, "weights": "random-init UNet (seeded He init, randomised BN statistics), 31.04 M params",
            "parallelism": f"clip-sharded dp{args.gpus}" if args.gpus > 1 else "single GPU"}


# ---------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from audiodenoiser_b200 import _lib, sharding, spectral, synth
    from audiodenoiser_b200.checkpoint import seeded_state_dict
    from audiodenoiser_b200.model import UNet
    from audiodenoiser_b200.pipeline import Denoiser, stream_host_batches

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    n_gpus = world
    _lib.check(_lib.load().adn_device_check(), "adn_device_check")

    batch = args.batch
    sr, length = synth.VARIANTS[args.variant]
    t_frames = spectral.num_frames(length, True)
    n_out = 128 * (t_frames - 1)
    peaks = measured_peaks()

    # ---- BASELINE config 2 (dataset-creation path): the STFT / iSTFT kernels alone on 4 096 clips, 3 s @ 8 kHz (the shape the
    # reference's create_*_dataset.py scripts feed, 393 MB in / 792 MB out: >> L2) -- the "STFT HBM GB/s" half of the metric.
    # Measured BEFORE the step loop: these kernels are bound by the fp32 issue port, i.e. by the SM clock, and the conv steps
    # below leave the chip at its power cap (1.45 GHz of 1.965) for a while -- each kernel is timed alone, as the roofline asks.
    kernels_c2 = {}
    if rank == 0 and world == 1 and args.c2_clips > 0:
        kernels_c2 = spectral_config2(args.c2_clips, dev, peaks, max(10, min(args.steps, 20)))

    sd = seeded_state_dict(7)
    net = UNet().eval()
    net.load_state_dict(sd)
    den = Denoiser(net, center=True, seed=1234 + rank, use_graph=True)
    job = sharding.ShardedDenoiser(den)

    # synthetic inputs: `rot` distinct device-resident batches so successive steps never re-read a cached input
    rot = args.rotate
    uniq = min(batch, 16)
    host_batches, clean_mags = [], []
    for r in range(rot):
        noisy, clean = [], []
        for i in range(uniq):
            a, b = synth.make_clip((rank * rot + r) * uniq + i, args.variant, return_clean=True)
            noisy.append(a); clean.append(b)
        idx = [i % uniq for i in range(batch)]
        gain = np.array([1.0 - 0.25 * ((i // uniq) % 3) / 3.0 for i in range(batch)], np.float32)[:, None]
        host_batches.append(torch.from_numpy(np.stack(noisy)[idx] * gain).pin_memory())
        clean_mags.append(spectral.stft_mag_batched(torch.from_numpy(np.stack(clean)[idx] * gain).to(dev), True))
    dev_batches = [h.to(dev) for h in host_batches]
    n_total = batch * n_gpus
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_device(i):
        audio, sums = job.step(dev_batches[i % rot], n_total, clean_mags[i % rot], gather=True, overlap_gather=True)
        return audio, sums

    host_outs = [torch.empty((batch, n_out), dtype=torch.float32).pin_memory() for _ in range(2)]
    host_stats = [torch.empty(8, dtype=torch.float64).pin_memory() for _ in range(2)]

    def run_host(first, count):
        """`count` end-to-end steps through the script-facing streaming entry: every step copies its batch from pinned host
        memory to the device, denoises, and copies the audio and the statistics back to pinned host memory; the copies of
        neighbouring steps overlap with compute (pipeline.stream_host_batches)."""
        def fn(i, wave_dev):
            # every rank reads ITS OWN rows and the group-reduced statistics back to the host; the all-gather of the audio runs on the
            # copy engines under the next step (as in the device-timed loop) and is completed inside the timed region by finish()
            _all, sums = job.step(wave_dev, n_total, clean_mags[(first + i) % rot], gather=True, overlap_gather=True)
            if world > 1 and i > 0:
                sums = job.wait_sums(previous=True)      # N > 1: step i reads back the reduced statistics of step i - 1 (ranks stay decoupled)
            return job.last_local, sums
        stream_host_batches(fn, [host_batches[(first + i) % rot] for i in range(count)], [host_outs[i & 1] for i in range(count)],
                            [host_stats[i & 1] for i in range(count)], dev)
        job.finish()
        if world > 1:                                    # ... and the last step's after the loop
            host_stats[count & 1].copy_(job.wait_sums(), non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def timed(fn, steps, warmup, sampler):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler:
            e0.record()
            for i in range(steps):
                fn(warmup + i)
            job.finish()                          # the last steps' overlapped all-gathers complete inside the timed region
            e1.record()
            barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    sampler = ClockSampler(local_rank)
    launches0 = net.launch_count
    ms_dev = timed(step_device, args.steps, args.warmup, sampler)
    clocks = sampler.summary()
    comm = None
    if world > 1:
        # the same steps with NO communication (no gather, local statistics only): the difference is what the collectives cost
        ms_nocomm = timed(lambda i: job.step(dev_batches[i % rot], n_total, clean_mags[i % rot], gather=False, reduce=False),
                          args.steps, args.warmup, ClockSampler(local_rank))
        comm = {"gather": job.gather_impl, "statistics": "ncclAllReduce of 8 float64, asynchronous",
                "ms_per_step_without_communication": ms_nocomm / args.steps,
                "comm_exposed_ms": (ms_dev - ms_nocomm) / args.steps}
    run_host(0, args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record()
    run_host(args.warmup, args.steps)
    e1.record()
    torch.cuda.synchronize()
    t_wall = (time.perf_counter() - t_wall) * 1e3                  # this rank's steps are done (streams drained); then the max over ranks
    barrier()
    e2e_event_ms, e2e_wall_ms = max_over_ranks(e0.elapsed_time(e1)), max_over_ranks(t_wall)
    ms_e2e = max(e2e_event_ms, e2e_wall_ms)
    audio, sums = step_device(0)
    job.finish()
    stats = sharding.stats_from_sums(sums.cpu())
    barrier()

    # ---- N > 1: the NCCL-gathered output must equal what ONE GPU computes for the same clips (VERDICT r1 item 1d / missing 8):
    # rank 0 regenerates rank 1's first batch from its seeds, denoises it with rank 1's phase seed and compares bit for bit.
    gather_check = None
    if world > 1:
        gathered = audio.clone()
        if rank == 0:
            other = 1
            noisy = [synth.make_clip((other * rot + 0) * uniq + i, args.variant) for i in range(uniq)]
            idx = [i % uniq for i in range(batch)]
            gain = np.array([1.0 - 0.25 * ((i // uniq) % 3) / 3.0 for i in range(batch)], np.float32)[:, None]
            wave1 = torch.from_numpy(np.stack(noisy)[idx] * gain).to(dev)
            k = den.last_seed - den.seed                          # every rank has made the same number of calls: call k used seed + k
            solo = Denoiser(net, center=True, seed=1234 + other + k, use_graph=False).denoise(wave1)
            rows = gathered[other * batch:(other + 1) * batch]
            own = Denoiser(net, center=True, seed=1234 + k, use_graph=False).denoise(dev_batches[0])
            gather_check = {"what": "all_gather rows of rank 1 (and rank 0's own rows) vs a single-GPU recomputation on rank 0",
                            "rows_checked": int(2 * batch), "bit_equal": bool(torch.equal(rows, solo) and torch.equal(gathered[:batch], own)),
                            "max_abs_diff": float(max((rows - solo).abs().max(), (gathered[:batch] - own).abs().max()))}
            del solo, own, wave1
        barrier()

    failed = False
    audio_s_per_step = n_total * CLIP_SECONDS
    value = audio_s_per_step * args.steps / (ms_dev * 1e-3)
    e2e_value = audio_s_per_step * args.steps / (ms_e2e * 1e-3)

    # ---- per-kernel figures, eager mode with CUDA events around every launch (same stream as the launches)
    kernels, roofline, launches_per_step = {}, None, None
    if True:
        net.profile = []
        eager = Denoiser(net, center=True, seed=1234 + rank, use_graph=False)
        reps = max(2, min(args.steps, 5))
        spans = {"stft": [], "istft": []}
        lc0 = net.launch_count
        for i in range(2 + reps):
            if i == 2:
                net.profile = []
                lc0 = net.launch_count
            w = dev_batches[i % rot]
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record(); mag = spectral.stft_mag_batched(w, True); ev[1].record()
            dmag = net(mag.unsqueeze(1)).squeeze(1)
            ev[2].record(); spectral.istft_batched(dmag, None, seed=i); ev[3].record()
            if i >= 2:
                spans["stft"].append((ev[0], ev[1])); spans["istft"].append((ev[2], ev[3]))
        torch.cuda.synchronize()
        launches_per_step = (net.launch_count - lc0) // reps + 2 + 1 + 3    # + stft + istft + error-sums + 3 loss kernels
        prof = net.profile
        net.profile = None
        agg = {}
        for layer, kind, flops, a, b in prof:
            d = agg.setdefault(kind, {"ms": 0.0, "flops": 0.0, "launches": 0})
            d["ms"] += a.elapsed_time(b); d["flops"] += flops; d["launches"] += 1
        per_layer = {}
        for layer, kind, flops, a, b in prof:
            d = per_layer.setdefault(layer, {"ms": 0.0, "flops": 0.0, "n": 0})
            d["ms"] += a.elapsed_time(b); d["flops"] += flops; d["n"] += 1
        tc = {"ms": 0.0, "flops": 0.0, "launches": 0}
        for kind in ("conv3x3", "conv3x3_head", "conv3x3_upm"):      # conv3x3_upm counts the MACs it EXECUTES (9 skip + 4 merged taps)
            if kind in agg:
                for k in tc:
                    tc[k] += agg[kind][k]
        tc_tflops = tc["flops"] / (tc["ms"] * 1e-3) / 1e12 if tc["ms"] > 0 else 0.0
        peak_tc = peaks["bf16_sustained"]
        roofline = {"kernel": "conv3x3_halo_kernel + conv3x3_dx_kernel + conv3x3_upm_kernel / conv3x3_upm2_kernel (tcgen05 implicit-GEMM Conv3x3+BN+ReLU: 10 + 3 + 4 launches/step; the 64-output-channel layers incl. the head-fused one run the kx-in-N variant; the first conv of decoder levels 1-3 runs with the ConvTranspose2d merged into its weights and is counted with the FLOPs it EXECUTES (9 skip + 4 merged taps), not with those of the two layers it replaces)", "bound": "tensor",
                    "achieved": tc_tflops, "peak": peak_tc, "unit": "TFLOP/s", "frac": tc_tflops / peak_tc,
                    "peak_source": f"MEASURED_PEAKS bf16_tflops_sustained ({peaks['source']}: cuBLAS 8192^3 back to back, power-capped); kernel timed inside a long step.  frac > 1 means these kernels hold higher clocks under the same power cap than the cuBLAS GEMM the peak was measured with",
                    "frac_of_burst_peak": tc_tflops / peaks["bf16_burst"], "burst_peak": peaks["bf16_burst"],
                    "launches_per_step": tc["launches"] // reps, "avg_launch_ms": tc["ms"] / max(tc["launches"], 1),
                    "algorithmic_flops_per_step": tc["flops"] / reps, "traffic": None}
        try:                                   # measured DRAM bytes per launch of the same kernel, from the committed ncu capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic.json"))).get(args.variant)
            if tr and tr["batch"] == batch:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_unit"] = "bytes per launch (dram read + write) from the committed ncu --set full capture profiles/conv_traffic.json -- NOT measured in this run"
                roofline["algorithmic_bytes_per_launch"] = unet_conv_bytes(batch, 257, t_frames) / max(tc["launches"] // reps, 1)
        except Exception:  # noqa: BLE001
            pass
        stft_ms = float(np.mean([a.elapsed_time(b) for a, b in spans["stft"]]))
        istft_ms = float(np.mean([a.elapsed_time(b) for a, b in spans["istft"]]))
        stft_bytes = batch * (4 * length + 4 * 257 * t_frames)
        istft_bytes = batch * (4 * 257 * t_frames + 4 * n_out)           # magnitude in, device-generated phase, audio out
        kernels = {
            "stft_mag": {"ms": stft_ms, "bytes": stft_bytes, "GBps": stft_bytes / stft_ms / 1e6, "frac_of_hbm": stft_bytes / stft_ms / 1e6 / peaks["hbm_gbs"]},
            "istft_ola": {"ms": istft_ms, "bytes": istft_bytes, "GBps": istft_bytes / istft_ms / 1e6, "frac_of_hbm": istft_bytes / istft_ms / 1e6 / peaks["hbm_gbs"]},
        }
        for kind, d in agg.items():
            kernels[kind] = {"ms_per_step": d["ms"] / reps, "launches_per_step": d["launches"] // reps,
                             "TFLOPs": (d["flops"] / (d["ms"] * 1e-3) / 1e12) if d["flops"] else None}
        if args.layers:
            kernels["layers"] = {k: {"ms": v["ms"] / v["n"], "TFLOPs": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] else None}
                                 for k, v in per_layer.items()}

    kernels.update(kernels_c2)

    # ---- parity on the benchmark configuration, in the run (rank 0, N = 1): clip 0 of timed batch 0
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        seed0 = 1234 + rank
        a0, m0, d0 = eager.denoise(dev_batches[0], return_spectrograms=True)
        ph0 = spectral.random_phasor(eager.last_seed, batch, t_frames, dev)[0].cpu().numpy()
        clean0 = synth.make_clip((rank * rot + 0) * uniq + 0, args.variant, return_clean=True)[1]
        parity = parity_on_benchmark_clip(host_batches[0][0].numpy(), clean0, sd, seed0, m0[0].cpu().numpy(), d0[0].cpu().numpy(),
                                          a0[0].cpu().numpy(), ph0)
        del a0, m0, d0

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload through the oracle chain
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        waves = [host_batches[0][i].numpy() for i in range(args.ref_clips)]
        cpu_reference_chain(waves[:1], sd, cores)                       # warm-up (thread pools, oneDNN primitives)
        dt = cpu_reference_chain(waves, sd, cores)
        cpu_baseline = {"value": len(waves) * CLIP_SECONDS / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{len(waves)} clip(s) of the same batch, variant {args.variant}: float64 numpy/scipy STFT -> fp32 torch-CPU "
                                  f"UNet -> {GL_ITERATIONS}-iteration istft/stft loop + istft; {dt:.2f} s"}

    gather_impl = job.gather_impl
    train = None
    if args.train_steps > 0:
        del den, job, eager
        net._ws = {}
        torch.cuda.empty_cache()
        train = run_train_bench(args, dev, world, rank, barrier, max_over_ranks)
        if args.train_batch != 64:            # SURVEY 8d config 5 names both batch 16 (train.py default) and 64 per GPU
            big = run_train_bench(args, dev, world, rank, barrier, max_over_ranks, batch_override=64)
            train["batch64"] = {k: big[k] for k in ("ms_per_step", "pairs_per_s", "e2e_pairs_per_s", "tflops")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {**workload_config(args, t_frames, batch),
                       "l2": f"inputs rotate over {rot} resident batches; every step rewrites ~{unet_workspace_gb(batch, 257, t_frames):.1f} GB of activations (>> 126 MB L2)",
                       "collectives": (f"all-gather of the audio ({gather_impl}), overlapped with the next step's kernels + asynchronous all_reduce(error sums) per step"
                                       if n_gpus > 1 else "none (single GPU)")},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "device_event_ms_per_step": e2e_event_ms / args.steps, "host_wall_ms_per_step": e2e_wall_ms / args.steps,
                    "h2d_bytes_per_step": int(batch * length * 4) * n_gpus, "d2h_bytes_per_step": int(batch * n_out * 4 + 32) * n_gpus},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "roofline": roofline,
            "kernels": kernels,
            "unet_flops_per_clip": unet_flops(257, t_frames),
            "unet_tflops_in_step": unet_flops(257, t_frames) * batch / (ms_dev / args.steps * 1e-3) / 1e12,
            "quality": {"snr_db_vs_clean_mag": stats["snr_db"], "l1_vs_clean_mag": stats["l1"],
                        "combined_perceptual_loss": {k: stats.get(k) for k in ("loss_total", "loss_stft", "loss_mel", "loss_l1")}, "note": "random-init weights: numbers only prove the statistics path runs"},
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "gather_check": gather_check,
            "communication": comm,
            "train_step": train,
        }
        emit(line)
        if parity is not None and not parity["pass"]:
            print(f"PARITY FAILURE on the benchmark configuration: {parity}", file=sys.stderr)
            failed = True
        if gather_check is not None and not gather_check["bit_equal"]:
            print(f"GATHER CHECK FAILURE: {gather_check}", file=sys.stderr)
            failed = True
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    return 1 if failed else 0


def spectral_config2(n_clips, dev, peaks, reps):
    """STFT / iSTFT kernel figures at BASELINE config 2: n_clips x 24 000 samples (center=True, create_test_dataset.py:39-40),
    n_clips x 16 000 (center=False, create_train_dataset.py:167-173), iSTFT with the device-generated phase and with an explicit
    phasor (test.py:36-37,40,48).  CUDA events on the launching stream around `reps` back-to-back launches after 3 warm-ups; every
    launch streams more than the 126 MB L2 holds.  Bytes are the algorithmic ones of SURVEY 8d."""
    import torch
    from audiodenoiser_b200 import spectral
    out = {}

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def entry(ms, nbytes, what):
        gbps = nbytes / ms / 1e6
        return {"ms": ms, "bytes": int(nbytes), "GBps": gbps, "frac_of_hbm": gbps / peaks["hbm_gbs"], "workload": what}

    g = torch.Generator(device=dev).manual_seed(0)
    for name, length, center in (("stft_mag_c2", 24000, True), ("stft_mag_train_c2", 16000, False)):
        x = torch.rand((n_clips, length), device=dev, generator=g) * 2 - 1
        t = spectral.num_frames(length, center)
        mag = torch.empty((n_clips, 257, t), device=dev)
        ms = timeit(lambda: spectral.stft_mag_batched(x, center, out=mag))
        out[name] = entry(ms, n_clips * (4 * length + 4 * 257 * t), f"{n_clips} clips x {length} samples, center={center} -> (257,{t})")
        if center:
            audio = torch.empty((n_clips, 128 * (t - 1)), device=dev)
            ms = timeit(lambda: spectral.istft_batched(mag, None, seed=1, out=audio))
            out["istft_seeded_c2"] = entry(ms, n_clips * (4 * 257 * t + 4 * 128 * (t - 1)), f"{n_clips} x (257,{t}) magnitudes, phase drawn in-kernel")
            ph = torch.polar(torch.ones_like(mag), torch.rand(mag.shape, device=dev, generator=g) * 6.2831853)
            ms = timeit(lambda: spectral.istft_batched(mag, ph, out=audio))
            out["istft_phasor_c2"] = entry(ms, n_clips * (12 * 257 * t + 4 * 128 * (t - 1)), f"{n_clips} x (257,{t}) magnitudes x explicit complex64 phasor")
            del ph, audio
        del x, mag
    torch.cuda.empty_cache()
    return out


def run_train_bench(args, dev, world, rank, barrier, max_over_ranks, batch_override=None):
    """BASELINE config 5: the train.py:65-72 step (train-mode forward + CombinedPerceptualLoss + backward + clip + AdamW) on
    synthetic (B,1,256,64) spectrogram pairs after the loader's float16 round trip, data-parallel over the ranks (one NCCL
    all-reduce of the flat 31 M-element gradient per step).  Returns the `train_step` object of the JSON line."""
    import numpy as np
    import torch
    from audiodenoiser_b200.checkpoint import seeded_state_dict
    from audiodenoiser_b200.model import UNet
    from audiodenoiser_b200.training import TrainEngine

    b = args.train_batch if batch_override is None else batch_override
    net = UNet()
    net.load_state_dict(seeded_state_dict(7))
    eng = TrainEngine(net, lr=1e-4, device=dev)
    g = torch.Generator().manual_seed(4321 + rank)
    pairs = []
    for _ in range(4):                                       # rotating resident batches (activations alone are >> L2)
        clean = (torch.rand((b, 1, 256, 64), generator=g) * 2.0).half().float()
        noisy = (clean + 0.3 * torch.rand((b, 1, 256, 64), generator=g)).half().float()
        pairs.append((noisy.pin_memory(), clean.pin_memory()))
    dev_pairs = [(a.to(dev), c.to(dev)) for a, c in pairs]
    steps, warm = args.train_steps, 3

    def timed(fn):
        for i in range(warm):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warm + i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    lc0 = eng.launch_count
    ms_dev = timed(lambda i: eng.train_step_graphed(*dev_pairs[i % 4]))
    launches = (eng.launch_count - lc0) // (steps + warm)

    def host_step(i):
        a, c = pairs[i % 4]
        losses = eng.train_step_graphed(a.to(dev, non_blocking=True), c.to(dev, non_blocking=True))
        return losses.cpu()                                  # loss.item() of train.py:72 (synchronises)
    ms_e2e = timed(host_step)
    losses = eng.train_step_graphed(*dev_pairs[0]).cpu().tolist()
    flops = 3.0 * unet_flops(256, 64) * b                    # forward + data gradient + weight gradient
    out = {"workload": f"BASELINE config 5: train.py step, batch {b} x (1,256,64) per GPU, AdamW lr 1e-4, clip 1.0, train-mode BatchNorm",
           "ms_per_step": ms_dev, "pairs_per_s": b * world / (ms_dev * 1e-3), "e2e_ms_per_step": ms_e2e,
           "e2e_pairs_per_s": b * world / (ms_e2e * 1e-3), "tflops": flops / (ms_dev * 1e-3) / 1e12, "kernel_launches_per_step": launches, "cuda_graph": True,
           "h2d_bytes_per_step": 2 * b * 256 * 64 * 4 * world, "d2h_bytes_per_step": 16 * world,
           "losses_after_warmup": losses, "parallelism": f"ddp{world}: all_reduce(AVG) of the flat fp32 gradient (31.04 M elements) in three buckets launched from backward (decoder 39 %, bottleneck 46 %, encoder 15 %), NCCL stream under the remaining backward kernels" if world > 1 else "single GPU"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and batch_override is None:
        from oracle.train_oracle import TrainOracle
        torch.set_num_threads(os.cpu_count() or 1)
        nb = min(b, 4)
        orc = TrainOracle(seeded_state_dict(7), lr=1e-4)
        a, c = pairs[0]
        orc.train_step(a[:nb], c[:nb])
        t0 = time.perf_counter()
        orc.train_step(a[:nb], c[:nb])
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": nb / dt, "unit": "pairs/s", "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": f"one step at batch {nb} of the same pairs: torch-CPU fp32 model.py (train mode) + loss.py + autograd + AdamW; {dt:.2f} s"}
    # the captured step holds NCCL kernels: it has to be gone before destroy_process_group (module <-> engine is a reference cycle)
    torch.cuda.synchronize()
    eng.release_graph()
    net._engine = None
    del eng, net
    import gc
    gc.collect()
    return out


def unet_conv_bytes(n, h, w):
    """Activation bytes (bf16, in + out, + pooled outputs) the 17 tensor-core conv launches of one forward must move."""
    hs, ws = [h], [w]
    for _ in range(4):
        hs.append(hs[-1] // 2); ws.append(ws[-1] // 2)
    ch = [64, 128, 256, 512, 1024]
    px = [n * hs[l] * ws[l] for l in range(5)]
    total = 0
    for l in range(5):
        cin = ch[l - 1] if l else None
        if l:
            total += px[l] * (cin + ch[l]) * 2                          # first conv of the level (level 0's is the Cin=1 direct conv)
        total += px[l] * (ch[l] + ch[l]) * 2                            # second conv
        if l < 4:
            total += px[l + 1] * ch[l] * 2                              # fused max-pool output
    for l in (3, 2, 1, 0):
        if l > 0:
            total += (px[l] * (ch[l] + ch[l]) + px[l + 1] * ch[l + 1]) * 2   # merged ConvTranspose + conv: skip and LOW tensor in, out
        else:
            total += px[l] * (2 * ch[l] + ch[l]) * 2                    # conv over [skip, up]
        total += px[l] * (ch[l] + (ch[l] if l else 0)) * 2              # second conv (level 0: fused head, fp32 out below)
    total += px[0] * 4
    return float(total)


def unet_workspace_gb(n, h, w):
    hs, ws = [h], [w]
    for _ in range(4):
        hs.append(hs[-1] // 2); ws.append(ws[-1] // 2)
    ch = [64, 128, 256, 512, 1024]
    total = 0
    for l in range(5):
        total += (2 if l == 4 else 4) * n * hs[l] * ws[l] * ch[l] * 2
        if l < 4:
            total += n * hs[l + 1] * ws[l + 1] * ch[l] * 2
    return total / 1e9


def _claim_stdout():
    """The driver parses ONE JSON line from stdout: keep a private handle on the real stdout and point fd 1 at stderr so
    that library chatter (e.g. "NCCL version ..." under NCCL_DEBUG=VERSION) cannot land in front of it."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def emit(line: dict):
    OUT.write(json.dumps(line) + "\n")
    OUT.flush()


OUT = sys.stdout


def main():
    global OUT
    OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="B", choices=["B", "R"], help="B: 3 s @ 44.1 kHz (BASELINE-literal); R: 3 s @ 8 kHz (what the reference scripts feed)")
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--rotate", type=int, default=4, help="distinct resident input batches")
    ap.add_argument("--ref-clips", type=int, default=1, help="clips per step of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c2-clips", type=int, default=4096, help="clips of the BASELINE config-2 STFT/iSTFT kernel leg (0 = skip)")
    ap.add_argument("--layers", action="store_true", help="add per-layer timings to the JSON line")
    ap.add_argument("--train-steps", type=int, default=10, help="timed train.py steps for the `train_step` object (0 = skip)")
    ap.add_argument("--train-batch", type=int, default=16, help="spectrogram pairs per GPU per training step (train.py default 16)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        os.dup2(OUT.fileno(), 1)                  # the child ranks print the line themselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())

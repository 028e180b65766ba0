"""Clip-level sharding of the hot path across the GPUs of one box (BASELINE config 4).

Every clip is independent through STFT -> eval-mode UNet -> iSTFT (model.py has no cross-sample op), so the path
shards with NO data-path collective: rank r owns the contiguous block ``shard_range(N, G, r)``.  NCCL is used only
after the fact, as BASELINE.json's north_star asks: one all-gather of the denoised outputs and one all-reduce(SUM) of
the partial error sums (so SNR / L1 are exact for unequal shards -- the reference's metrics are full-batch means,
loss.py:86).  The helpers take a ``torch.distributed`` process group and work on whatever device the tensors live on,
which lets the host logic be tested with the gloo backend on CPU (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block of ceil(N/G) items for ``rank`` (the last ranks may get fewer, or none)."""
    if world_size < 1 or not (0 <= rank < world_size) or n_items < 0:
        raise ValueError("bad shard request")
    per = -(-n_items // world_size)
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def shard_sizes(n_items: int, world_size: int) -> list[int]:
    return [hi - lo for lo, hi in (shard_range(n_items, world_size, r) for r in range(world_size))]


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def all_gather_rows(local: torch.Tensor, n_total: int, group=None, out: torch.Tensor | None = None, async_op: bool = False):
    """Gather the row-sharded ``local`` (n_r, ...) of every rank into (n_total, ...) on every rank, in rank order.

    Shards follow ``shard_range``: equal blocks of ceil(N/G) rows except the tail.  One ``all_gather_into_tensor`` on a
    padded (G*per, ...) buffer; the padding rows of the tail ranks are dropped by the final slice (a view).
    ``async_op=True`` returns ``(gathered, work)``: the local shard is first copied into its slot of ``out`` on the current
    stream (so ``local`` may be overwritten right away) and the collective runs on the backend's stream; ``work.wait()``
    before reading ``gathered``."""
    world, rank = _world(group)
    if world == 1:
        if local.shape[0] != n_total:
            raise ValueError("single-rank gather: local rows != n_total")
        return local
    per = -(-n_total // world)
    lo, hi = shard_range(n_total, world, rank)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: expected {hi - lo} local rows, got {local.shape[0]}")
    tail = tuple(local.shape[1:])
    if out is None:
        out = torch.empty((world * per,) + tail, dtype=local.dtype, device=local.device)
    elif tuple(out.shape) != (world * per,) + tail or out.dtype != local.dtype:
        raise ValueError("out must be (world*ceil(N/world), ...) of the local dtype")
    if async_op:
        slot = out[rank * per:(rank + 1) * per]
        slot[: local.shape[0]].copy_(local)
        if local.shape[0] < per:
            slot[local.shape[0]:].zero_()
        work = dist.all_gather_into_tensor(out, slot, group=group, async_op=True)
        return out[:n_total], work
    if local.shape[0] == per:
        src = local.contiguous()
    else:                                   # tail rank: pad to the common block size
        src = torch.zeros((per,) + tail, dtype=local.dtype, device=local.device)
        src[: local.shape[0]].copy_(local)
    dist.all_gather_into_tensor(out, src, group=group)
    return out[:n_total]


class PeerGather:
    """All-gather of the row-sharded audio over NVLink by the COPY ENGINES instead of ncclAllGather.

    Why: the hot path's kernels are persistent (one CTA per SM, tiles assigned statically); an NCCL kernel that holds even a
    few SMs while they run delays the CTAs that land on those SMs by a whole kernel (round 1: 1 -> 8 GPUs lost 5 %, most of it
    here).  A peer copy needs no SM.  Every rank owns a symmetric (world * per, ...) buffer (torch symmetric memory: cuMem
    allocations mapped into every peer over NVLink / NVSwitch); a step copies the local shard into slot `rank` of EVERY rank's
    buffer with cudaMemcpyAsync on a side stream, then joins a device-side barrier of the group on the same stream.  After the
    barrier the local buffer holds all shards in rank (= clip) order.  Two buffers alternate, so the gather of step i overlaps the
    kernels of step i + 1.  `available()` is False without CUDA symmetric memory (then the caller uses NCCL)."""

    def __init__(self, per: int, tail: tuple, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = _world(group)
        self.per, self.tail = int(per), tuple(tail)
        grp = group if group is not None else dist.group.WORLD
        self.stream = torch.cuda.Stream(device)
        self.bufs, self.handles, self.peers = [], [], []
        for _ in range(2):
            buf = symm.empty((self.world * self.per,) + self.tail, dtype=dtype, device=device)
            hdl = symm.rendezvous(buf, grp)
            self.bufs.append(buf)
            self.handles.append(hdl)
            self.peers.append([hdl.get_buffer(r, tuple(buf.shape), dtype) for r in range(self.world)])
        self.turn = 0
        self.pending = [None, None]

    @staticmethod
    def available(device) -> bool:
        try:
            import torch.distributed._symmetric_memory as symm  # noqa: F401
        except Exception:  # noqa: BLE001
            return False
        return torch.device(device).type == "cuda" and dist.is_initialized() and dist.get_backend() == "nccl"

    def start(self, local: torch.Tensor, n_total: int):
        """Begin gathering `local` (n_r rows) of this step; returns the (n_total, ...) view of the buffer it lands in (complete
        after `finish()`, or once the event of this turn has been waited on)."""
        b = self.turn
        self.turn ^= 1
        comp = torch.cuda.current_stream(local.device)
        if self.pending[b] is not None:                        # the gather that last used this buffer (two steps ago)
            comp.wait_event(self.pending[b])
        lo = self.rank * self.per
        own = self.bufs[b][lo:lo + local.shape[0]]
        own.copy_(local, non_blocking=True)                    # own slot first, on the compute stream: `local` may be reused at once
        ready = torch.cuda.Event()
        ready.record(comp)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for r in range(1, self.world):                     # ring order: at any moment every rank receives from one peer
                peer = (self.rank + r) % self.world
                self.peers[b][peer][lo:lo + local.shape[0]].copy_(own, non_blocking=True)
            # device-side barrier of the group on this stream: when it completes here, every rank's copies into THIS rank's buffer
            # have landed.  It also orders the steps across ranks: nobody can start step i + 2 (same buffer) before all finished step i.
            self.handles[b].barrier(channel=0)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.pending[b] = done
        return self.bufs[b][:n_total]

    def finish(self):
        comp = torch.cuda.current_stream()
        for i, ev in enumerate(self.pending):
            if ev is not None:
                comp.wait_event(ev)
                self.pending[i] = None


def all_reduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce(SUM) of a small float64 vector of partial sums / counts."""
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def stats_from_sums(sums) -> dict:
    """[sum|target-pred|, sum target^2, sum (target-pred)^2, count, n_clips, n*stft, n*mel, n*l1] -> metrics.
    SNR = 10 log10(sum target^2 / sum err^2) on magnitude spectrograms (the reference defines none); the loss
    terms are the clip-weighted means of CombinedPerceptualLoss (loss.py:83-95), exact for unequal shards."""
    s_abs, s_sig, s_err, cnt = (float(v) for v in sums[:4])
    l1 = s_abs / cnt if cnt > 0 else float("nan")
    snr = 10.0 * math.log10(s_sig / s_err) if s_err > 0 and s_sig > 0 else float("inf")
    out = {"l1": l1, "snr_db": snr, "count": int(cnt)}
    if len(sums) >= 8 and float(sums[4]) > 0:
        n = float(sums[4])
        stft, mel, l1m = float(sums[5]) / n, float(sums[6]) / n, float(sums[7]) / n
        out.update({"loss_total": 0.4 * stft + 0.4 * mel + 0.2 * l1m, "loss_stft": stft, "loss_mel": mel, "loss_l1": l1m,
                    "clips": int(n)})
    return out


class ShardedDenoiser:
    """One rank's view of the clip-sharded job: denoise the local shard with the single-GPU ``Denoiser``, then gather
    outputs and reduce the error statistics over the group."""

    def __init__(self, denoiser, group=None, with_loss: bool = True):
        self.denoiser = denoiser
        self.with_loss = bool(with_loss)
        if self.with_loss:
            from .loss import CombinedPerceptualLoss
            self._criterion = CombinedPerceptualLoss()
        self.group = group
        self.world, self.rank = _world(group)
        self._gather_buf = None
        self._gather_bufs = [None, None]
        self._gather_work = [None, None]
        self._gather_turn = 0
        self._sums = None
        self._sums_bufs = None
        self._sums_turn = 0
        self._sums_work = [None, None]
        self._peer = None                 # PeerGather (copy-engine all-gather), created on first use when available
        self._peer_failed = False
        self.last_local = None
        self.gather_impl = "none"

    def local_range(self, n_total: int) -> tuple[int, int]:
        return shard_range(n_total, self.world, self.rank)

    def error_sums(self, pred_mag: torch.Tensor, target_mag: torch.Tensor) -> torch.Tensor:
        """Device float64 [sum|d|, sum t^2, sum d^2, count] of this rank's shard (CUDA kernel adn_spec_error_sums_f64).  Two
        result buffers alternate, so the asynchronous all-reduce of one step never meets the next step's writes."""
        from . import _lib
        if not pred_mag.is_cuda:
            raise _lib.AdnError("error_sums needs CUDA tensors (no CPU fallback)")
        if self._sums_bufs is None or self._sums_bufs[0].device != pred_mag.device:
            self._sums_bufs = [torch.zeros(8, dtype=torch.float64, device=pred_mag.device) for _ in range(2)]
        self._sums_turn ^= 1
        if self._sums_work[self._sums_turn] is not None:      # the all-reduce that last used this buffer (two steps ago)
            self._sums_work[self._sums_turn].wait()
            self._sums_work[self._sums_turn] = None
        self._sums = self._sums_bufs[self._sums_turn]
        with torch.cuda.device(pred_mag.device):
            _lib.check(_lib.load().adn_zero_bytes(self._sums.data_ptr(), 64, _lib.stream_ptr()), "adn_zero_bytes")
        p = pred_mag.float().contiguous(); t = target_mag.float().contiguous()
        if p.shape != t.shape:
            raise ValueError("pred and target must have the same shape")
        with torch.cuda.device(p.device):
            st = _lib.load().adn_spec_error_sums_f64(p.data_ptr(), t.data_ptr(), p.numel(), self._sums.data_ptr(), _lib.stream_ptr())
        _lib.check(st, "adn_spec_error_sums_f64")
        loss4 = None
        if self.with_loss and p.shape[-1] > 31:
            # CombinedPerceptualLoss of this shard (test.py:118-122), weighted by its clip count so the reduced value is the
            # full-batch mean the reference would print
            loss4 = self._criterion.values(p.unsqueeze(1) if p.dim() == 3 else p, t.unsqueeze(1) if t.dim() == 3 else t)
        with torch.cuda.device(p.device):
            st = _lib.load().adn_stats_pack_f64(self._sums.data_ptr(), p.numel(), p.shape[0], loss4.data_ptr() if loss4 is not None else 0,
                                                _lib.stream_ptr())
        _lib.check(st, "adn_stats_pack_f64")
        return self._sums

    def wait_sums(self, previous: bool = False):
        """Make the CURRENT STREAM wait for the statistics all-reduce started by the latest ``step(..., overlap_gather=True)`` -- or,
        with ``previous=True``, by the step before it -- and return that step's (now reduced) statistics vector; no host
        synchronisation.  Reading every step's statistics one step late keeps the ranks decoupled: waiting for the CURRENT
        step's all-reduce makes every rank's stream wait for the slowest rank in every step."""
        turn = self._sums_turn ^ (1 if previous else 0)
        w = self._sums_work[turn]
        if w is not None:
            w.wait()
            self._sums_work[turn] = None
        return self._sums_bufs[turn] if self._sums_bufs is not None else None

    def finish(self):
        """Wait for the all-gathers started by ``step(..., overlap_gather=True)``."""
        if self._peer is not None:
            self._peer.finish()
        for i, w in enumerate(self._sums_work):
            if w is not None:
                w.wait()
                self._sums_work[i] = None
        for i, w in enumerate(self._gather_work):
            if w is not None:
                w.wait()
                self._gather_work[i] = None

    def step(self, wave_local: torch.Tensor, n_total: int, target_mag_local: torch.Tensor | None = None, gather: bool = True,
             overlap_gather: bool = False, reduce: bool = True):
        """Denoise this rank's clips; returns (audio, sums): ``audio`` is the gathered (n_total, samples) tensor when
        ``gather`` (every rank gets it, rank order = clip order) else the local shard; ``sums`` is the group-reduced
        float64 statistics vector (None without a target).

        ``overlap_gather=True``: the all-gather of this step's audio is started asynchronously into one of two alternating
        buffers and overlaps the NEXT step's kernels (nothing on the path depends on it); the returned ``audio`` of a step is
        complete once the step after next has been issued, or after ``finish()``."""
        if target_mag_local is not None:
            audio, _mag, den = self.denoiser.denoise(wave_local, return_spectrograms=True)
            sums = self.error_sums(den, target_mag_local)
            if not reduce:
                pass                                            # local partial sums only (bench: the no-communication reference run)
            elif self.world > 1 and overlap_gather and sums.is_cuda:
                # statistics: NCCL all-reduce started asynchronously (it runs on NCCL's stream); `sums` is final after finish()
                self._sums_work[self._sums_turn] = dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            else:
                sums = all_reduce_sums(sums, self.group)
        else:
            audio, sums = self.denoiser.denoise(wave_local), None
        self.last_local = audio                     # this rank's own rows (valid on the compute stream now, whatever the gather does)
        if gather and self.world > 1 and overlap_gather and not self._peer_failed and audio.is_cuda and os.environ.get("ADN_GATHER", "peer") != "nccl":
            per = -(-n_total // self.world)
            if self._peer is None and PeerGather.available(audio.device):
                try:
                    self._peer = PeerGather(per, tuple(audio.shape[1:]), audio.dtype, audio.device, self.group)
                    self.gather_impl = "copy-engine peer copies into symmetric buffers (no SM), device barrier per step"
                except Exception as exc:  # noqa: BLE001  (no symmetric memory on this box / torch build: use NCCL)
                    self._peer_failed = True
                    self.gather_impl = f"ncclAllGather (symmetric memory unavailable: {type(exc).__name__})"
            elif self._peer is None:
                self._peer_failed = True
        if gather and self.world > 1 and overlap_gather and self._peer is not None:
            audio = self._peer.start(audio, n_total)
        elif gather and self.world > 1 and overlap_gather:
            per = -(-n_total // self.world)
            if self.gather_impl == "none":
                self.gather_impl = "ncclAllGather on NCCL's stream"
            shape = (self.world * per,) + tuple(audio.shape[1:])
            b = self._gather_turn
            self._gather_turn ^= 1
            if self._gather_work[b] is not None:            # the gather that last used this buffer
                self._gather_work[b].wait()
            if self._gather_bufs[b] is None or tuple(self._gather_bufs[b].shape) != shape or self._gather_bufs[b].device != audio.device:
                self._gather_bufs[b] = torch.empty(shape, dtype=audio.dtype, device=audio.device)
            audio, self._gather_work[b] = all_gather_rows(audio, n_total, self.group, out=self._gather_bufs[b], async_op=True)
        elif gather and self.world > 1:
            per = -(-n_total // self.world)
            shape = (self.world * per,) + tuple(audio.shape[1:])
            if self._gather_buf is None or tuple(self._gather_buf.shape) != shape or self._gather_buf.device != audio.device:
                self._gather_buf = torch.empty(shape, dtype=audio.dtype, device=audio.device)
            audio = all_gather_rows(audio, n_total, self.group, out=self._gather_buf)
        return audio, sums


def average_gradients_(flat_grad, group=None):
    """DDP gradient averaging for the training step (BASELINE config 5): ONE all-reduce over the flat fp32 gradient buffer,
    in place.  torch DDP semantics: every rank back-propagates the mean loss of its own shard (BatchNorm statistics per
    replica, the reference has no SyncBN), and the gradients are averaged over the ranks.  NCCL averages in the collective;
    gloo (the CPU tests) sums and divides."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return flat_grad
    world = dist.get_world_size(group)
    if world == 1:
        return flat_grad
    if flat_grad.is_cuda:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.div_(world)
    return flat_grad


class _GradBucket:
    """One asynchronous all-reduce(AVG) over a slice of the flat gradient buffer (see TrainEngine.backward)."""

    def __init__(self, flat_slice, group):
        self.t, self.world, self.work = flat_slice, 1, None
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(group)
        if self.world > 1 and flat_slice.numel() > 0:
            op = dist.ReduceOp.AVG if flat_slice.is_cuda else dist.ReduceOp.SUM
            self.work = dist.all_reduce(flat_slice, op=op, group=group, async_op=True)

    def wait(self):
        if self.work is not None:
            self.work.wait()
            if not self.t.is_cuda:
                self.t.div_(self.world)
            self.work = None


def average_gradients_async(flat_slice, group=None) -> _GradBucket:
    """Start averaging one BUCKET of the flat gradient buffer over the ranks and return a handle (``.wait()``): the training engine
    launches a bucket as soon as the backward pass has produced its last gradient, so the collective runs on NCCL's stream
    under the remaining backward kernels (torch DDP's bucketed overlap)."""
    return _GradBucket(flat_slice, group)

"""The training step of the reference (code/train.py:65-72) on the B200 kernels of libadn_b200.so.

    optimizer.zero_grad(); outputs = model(noisy); loss, *_ = criterion(outputs, clean); loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0); optimizer.step()

``TrainEngine`` owns everything that body touches: fp32 master parameters (the drop-in ``UNet``'s own parameters are re-pointed
at slices of ONE flat buffer, so ``state_dict()`` keeps working and AdamW / the gradient norm / the DDP all-reduce each run over
one contiguous tensor), bf16 packed weights for the forward and data-gradient convolutions, the saved activations, and the
AdamW moments.  model.py is executed in train() mode: BatchNorm uses batch statistics and updates its running estimates
(model.py:12,15), exactly as ``nn.BatchNorm2d`` does.

Data-parallel training (BASELINE config 5): one process per GPU, every rank runs ``train_step`` on its shard of the batch and
the flat gradient buffer is averaged with ONE NCCL all-reduce between backward and the clip (torch DDP semantics: BatchNorm
statistics stay per replica, the reference has no SyncBN).

Numerics: bf16 activations / activation gradients / GEMM operands, fp32 accumulation, fp32 statistics, parameters, parameter
gradients and optimizer state.  Conv biases that feed a train-mode BatchNorm have a mathematically zero gradient (the batch
mean removes them); autograd produces float noise there, this engine writes exact zeros.
"""
from __future__ import annotations

from collections import OrderedDict

import os

import torch

from . import _lib
from .checkpoint import BN_EPS, BN_MOMENTUM, state_dict_spec
from .loss import mel_filterbank

_CH = [64, 128, 256, 512, 1024]
_BUFFER_LEAVES = ("running_mean", "running_var", "num_batches_tracked")


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


def flat_layout():
    """{key: (offset, numel, shape)} of every trainable tensor in state_dict order, and the padded total (in floats): the
    layout of the engine's flat parameter / gradient / moment buffers.  Slices start on 64-float (256-byte) boundaries."""
    offsets: "OrderedDict[str, tuple]" = OrderedDict()
    off = 0
    for key, (shape, _dt) in state_dict_spec().items():
        if key.rsplit(".", 1)[1] in _BUFFER_LEAVES:
            continue
        numel = 1
        for s in shape:
            numel *= s
        offsets[key] = (off, numel, tuple(shape))
        off += _align(numel)
    return offsets, off


class TrainEngine:
    """train.py:65-72 for ``audiodenoiser_b200.model.UNet``; AdamW defaults are torch's (train.py:124 passes only lr)."""

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 max_norm: float = 1.0, device=None, process_group=None):
        torch_ = _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch_.device("cuda", torch_.cuda.current_device()) if device is None else torch_.device(device)
        self.model = model.to(self.device)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = float(lr), tuple(float(b) for b in betas), float(eps), float(weight_decay), float(max_norm)
        self.group = process_group
        self.step_count = 0
        self.launch_count = 0
        dev = self.device

        # ---- flat fp32 parameter / gradient / moment buffers in state_dict order; 256-byte aligned slices
        self.offsets, off = flat_layout()
        self.numel = off
        self.P = torch.zeros(off, dtype=torch.float32, device=dev)
        self.G = torch.zeros(off, dtype=torch.float32, device=dev)
        self.M = torch.zeros(off, dtype=torch.float32, device=dev)
        self.V = torch.zeros(off, dtype=torch.float32, device=dev)
        params = dict(self.model.named_parameters())
        with torch.no_grad():
            for key, (o, n, shape) in self.offsets.items():
                view = self.P[o:o + n].view(shape)
                view.copy_(params[key].detach().to(dev, torch.float32))
                params[key].data = view                      # the module's parameters now alias the flat buffer
        self.buffers = dict(self.model.named_buffers())
        # the num_batches_tracked counters alias ONE int64 vector (like the parameters alias the flat buffer): one launch advances all
        nbt = [b for k, b in self.buffers.items() if k.endswith("num_batches_tracked")]
        self._nbt_flat = torch.zeros(len(nbt), dtype=torch.int64, device=dev)
        for i, b in enumerate(nbt):
            self._nbt_flat[i] = b.to(dev)
            b.data = self._nbt_flat[i]
        self.ones = torch.ones(1024, dtype=torch.float32, device=dev)
        self.zeros = torch.zeros(1024, dtype=torch.float32, device=dev)
        self.ws = torch.empty(int(self.lib.adn_train_workspace_bytes()), dtype=torch.uint8, device=dev)
        self.wg_ws = torch.empty(int(self.lib.adn_wgrad_workspace_bytes()), dtype=torch.uint8, device=dev)
        self.wg_ws2 = None                       # second split-K workspace: weight gradients alternate between two side streams
        self.norm_coef = torch.zeros(2, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.float32, device=dev)      # AdamW step count, advanced on the device
        self._graph = None
        self._bucketed = False            # backward() launches the DDP gradient buckets (set by train_step only)
        self._buckets = []
        self.mel_fb = mel_filterbank().to(dev)
        self.loss_out = torch.empty(4, dtype=torch.float32, device=dev)
        self._loss_ws = None
        self._wg_stream = None
        self._aux_stream = None
        self._dgrad_stale = False
        # ADN_SPLIT_PACK=1: data-gradient weight packs on a side stream under the forward instead of behind AdamW.  Measured SLOWER
        # (3.62 -> 3.74 ms at batch 16: the fp32 weights are read twice and the pack competes with the latency-bound forward kernels): off
        self.split_pack = os.environ.get("ADN_SPLIT_PACK", "0") == "1"
        self.overlap_wgrad = os.environ.get("ADN_OVERLAP_WGRAD", "1") != "0"      # weight gradients on a side stream (backward())

        # ---- layer table: (key prefix, conv idx, bn idx, level, c0, c1, co)
        self.layers = []
        cin = 1
        for l, name in enumerate(("downconv1", "downconv2", "downconv3", "downconv4")):
            self.layers.append((f"{name}.conv", 0, 1, l, cin, 0, _CH[l]))
            self.layers.append((f"{name}.conv", 3, 4, l, _CH[l], 0, _CH[l]))
            cin = _CH[l]
        self.layers.append(("bottleneck", 0, 1, 4, 512, 0, 1024))
        self.layers.append(("bottleneck", 3, 4, 4, 1024, 0, 1024))
        for i, l in enumerate((3, 2, 1, 0)):
            self.layers.append((f"upconv{i + 1}.conv", 0, 1, l, _CH[l], _CH[l], _CH[l]))
            self.layers.append((f"upconv{i + 1}.conv", 3, 4, l, _CH[l], 0, _CH[l]))
        self.stats = {}
        for (p, ci_, bi, l, c0, c1, co) in self.layers:
            self.stats[(p, ci_)] = tuple(torch.empty(co, dtype=torch.float32, device=dev) for _ in range(4))   # scale, shift, mean, invstd
        self.packed = {}
        self._pack_table = None
        self._packed_versions = None
        self._alloc_packed()
        self.repack()
        self.saved = None
        self.model._engine = self

    def set_hyperparameters(self, lr=None, betas=None, eps=None, weight_decay=None, max_norm=None):
        before = (self.lr, self.betas, self.eps, self.weight_decay, self.max_norm)
        if lr is not None: self.lr = float(lr)
        if betas is not None: self.betas = tuple(float(b) for b in betas)
        if eps is not None: self.eps = float(eps)
        if weight_decay is not None: self.weight_decay = float(weight_decay)
        if max_norm is not None: self.max_norm = float(max_norm)
        if (self.lr, self.betas, self.eps, self.weight_decay, self.max_norm) != before:
            self._graph = None                   # the values are baked into a captured step; unchanged values keep the graph

    def _versions(self):
        return tuple(int(p._version) for p in self.model.parameters())

    def refresh_if_parameters_changed(self):
        """A torch optimizer (or load_state_dict) updates the parameters in place and bumps their version counters; the bf16
        operand copies are then stale.  (The engine's own optimizer_step repacks by itself.)"""
        v = self._versions()
        if v != self._packed_versions:
            self.repack()

    # ------------------------------------------------------------------ views / packing
    def pview(self, key):
        o, n, shape = self.offsets[key]
        return self.P[o:o + n].view(shape)

    def gview(self, key):
        o, n, shape = self.offsets[key]
        return self.G[o:o + n].view(shape)

    def _pptr(self, key):
        return self.P.data_ptr() + 4 * self.offsets[key][0]

    def _gptr(self, key):
        return self.G.data_ptr() + 4 * self.offsets[key][0]

    def _alloc_packed(self):
        bf = dict(dtype=torch.bfloat16, device=self.device)
        for (p, ci_, bi, l, c0, c1, co) in self.layers:
            cin = c0 + c1
            if cin == 1:
                continue
            self.packed[(p, ci_)] = (torch.empty((co, 9, cin), **bf), torch.empty((cin, 9, co), **bf))     # forward, data-gradient
        for i, l in enumerate((3, 2, 1, 0)):
            ci, co = _CH[l + 1], _CH[l]
            self.packed[f"upconv{i + 1}.up"] = (torch.empty((4, co, ci), **bf), torch.empty((ci, 4 * co), **bf))

    def _build_pack_table(self):
        import numpy as np
        rec = np.dtype([("w", np.uint64), ("fwd", np.uint64), ("dgrad", np.uint64), ("c_out", np.int32), ("c_in", np.int32),
                        ("kind", np.int32), ("pad", np.int32)])
        rows = []
        for (p, ci_, bi, l, c0, c1, co) in self.layers:
            cin = c0 + c1
            if cin == 1:
                continue
            wf, wd = self.packed[(p, ci_)]
            rows.append((self._pptr(f"{p}.double_conv.{ci_}.weight"), wf.data_ptr(), wd.data_ptr(), co, cin, 0, (co // 32) * (cin // 32)))
        for i, l in enumerate((3, 2, 1, 0)):
            ci, co = _CH[l + 1], _CH[l]
            wf, wd = self.packed[f"upconv{i + 1}.up"]
            rows.append((self._pptr(f"upconv{i + 1}.up.weight"), wf.data_ptr(), wd.data_ptr(), co, ci, 1, max(1, (4 * co * ci) // (256 * 16))))
        host = np.array(rows, dtype=rec)
        self._pack_n = len(rows)
        self._pack_blocks = int(sum(r[6] for r in rows))          # `pad` = blocks per entry of the one-dimensional pack grid
        self._pack_table = torch.from_numpy(host.view(np.uint8).copy()).to(self.device)

    def repack(self, which: int = 3):
        """fp32 master weights -> bf16 GEMM operands (forward [Co][tap][Ci], data gradient [Ci][8-tap][Co]; convT likewise): one
        launch over a device-side table of all 21 conv / convT weights.  which: 1 = forward operands, 2 = data-gradient operands,
        3 = both.  train_step packs the forward operands after AdamW and the data-gradient operands on a side stream under the next
        forward (they are first read in backward)."""
        if self._pack_table is None:
            self._build_pack_table()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.adn_pack_weights_table_flat_bf16(self._pack_table.data_ptr(), self._pack_n, self._pack_blocks, which,
                                                                 _lib.stream_ptr()), "pack weights")
        self.launch_count += 1
        if which & 1:
            self.model._packed = None          # the eval-mode forward of the module repacks from the updated parameters
            self._packed_versions = self._versions()
        if which & 2:
            self._dgrad_stale = False
        else:
            self._dgrad_stale = True

    # ------------------------------------------------------------------ forward (model.py:70-94 in train() mode)
    def forward(self, x):
        if not x.is_cuda or x.dim() != 4 or x.shape[1] != 1:
            raise ValueError("expected a CUDA (N, 1, F, T) tensor")
        n, _, h, w = x.shape
        if h % 16 or w % 16:
            raise ValueError("the training path needs F and T divisible by 16 (SpectrogramDataset yields (256, 64), data_loader.py:9)")
        x = x.float().contiguous()
        lib, s = self.lib, _lib.stream_ptr()
        bf = dict(dtype=torch.bfloat16, device=self.device)
        hs = [h >> l for l in range(5)]
        wz = [w >> l for l in range(5)]
        sv = {"x": x, "n": n, "hs": hs, "wz": wz}

        def conv_bn_relu(layer, src0, src1):
            p, ci_, bi, l, c0, c1, co = layer
            hh, ww = hs[l], wz[l]
            pixels = n * hh * ww
            z = torch.empty((n, hh, ww, co), **bf)
            bias = self._pptr(f"{p}.double_conv.{ci_}.bias")
            if c0 + c1 == 1:
                st = lib.adn_conv3x3_c1_affine_bf16(x.data_ptr(), n, hh, ww, self._pptr(f"{p}.double_conv.{ci_}.weight"), self.ones.data_ptr(),
                                                    bias, 0, z.data_ptr(), s)
            else:
                st = lib.adn_conv3x3_affine_bf16(src0.data_ptr(), c0, src1.data_ptr() if src1 is not None else 0, c1, hh if src1 is not None else 0,
                                                 ww if src1 is not None else 0, n, hh, ww, self.packed[(p, ci_)][0].data_ptr(), co,
                                                 self.ones.data_ptr(), bias, 0, z.data_ptr(), s)
            _lib.check(st, f"conv fwd {p}.{ci_}")
            scale, shift, mean, invstd = self.stats[(p, ci_)]
            bn = f"{p}.double_conv.{bi}"
            _lib.check(lib.adn_bn_train_stats_f32(z.data_ptr(), pixels, co, self._pptr(f"{bn}.weight"), self._pptr(f"{bn}.bias"), BN_EPS, BN_MOMENTUM,
                                                  self.buffers[f"{bn}.running_mean"].data_ptr(), self.buffers[f"{bn}.running_var"].data_ptr(),
                                                  scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), self.ws.data_ptr(), s),
                       f"bn stats {bn}")
            y = torch.empty_like(z)
            _lib.check(lib.adn_bn_relu_apply_bf16(z.data_ptr(), scale.data_ptr(), shift.data_ptr(), pixels, co, y.data_ptr(), s), f"bn apply {bn}")
            self.launch_count += 4
            sv[(p, ci_)] = (z, y, src0, src1)
            return y

        with torch.cuda.device(self.device):
            li = iter(self.layers)
            cur = None
            skips = []
            for l in range(4):
                a = conv_bn_relu(next(li), cur, None)
                sk = conv_bn_relu(next(li), a, None)
                skips.append(sk)
                cur = torch.empty((n, hs[l + 1], wz[l + 1], _CH[l]), **bf)
                _lib.check(lib.adn_maxpool2x2_bf16(sk.data_ptr(), n, hs[l], wz[l], _CH[l], cur.data_ptr(), s), "maxpool")
                self.launch_count += 1
                sv[("pool", l)] = cur
            a = conv_bn_relu(next(li), cur, None)
            cur = conv_bn_relu(next(li), a, None)
            for i, l in enumerate((3, 2, 1, 0)):
                name = f"upconv{i + 1}.up"
                ci, co = _CH[l + 1], _CH[l]
                up = torch.empty((n, hs[l], wz[l], co), **bf)
                _lib.check(lib.adn_convt2x2_bf16(cur.data_ptr(), ci, n, hs[l + 1], wz[l + 1], self.packed[name][0].data_ptr(), co,
                                                 self._pptr(f"{name}.bias"), up.data_ptr(), s), f"convT {name}")
                self.launch_count += 1
                sv[name] = (cur, up)
                a = conv_bn_relu(next(li), skips[l], up)
                cur = conv_bn_relu(next(li), a, None)
            out = torch.empty((n, 1, h, w), dtype=torch.float32, device=self.device)
            _lib.check(lib.adn_head1x1_forward_f32(cur.data_ptr(), self._pptr("out.weight"), self._pptr("out.bias"), n * h * w, out.data_ptr(), s), "head")
            self.launch_count += 1
            sv["head_in"] = cur
            _lib.check(lib.adn_i64_add_n(self._nbt_flat.data_ptr(), self._nbt_flat.numel(), 1, s), "nbt")      # the 18 num_batches_tracked counters
            self.launch_count += 1
        self.saved = sv
        return out

    # ------------------------------------------------------------------ backward (loss.backward(), train.py:69)
    def backward(self, d_out):
        """Adds the parameter gradients of sum(d_out * model(x)) into the flat gradient buffer (zero it with zero_grad())."""
        sv = self.saved
        if sv is None:
            raise RuntimeError("backward() needs a forward() first")
        lib, s = self.lib, _lib.stream_ptr()
        n, hs, wz = sv["n"], sv["hs"], sv["wz"]
        bf = dict(dtype=torch.bfloat16, device=self.device)
        d_out = d_out.float().contiguous()
        wsp = self.ws.data_ptr()
        if self._dgrad_stale:                  # a backward outside train_step after a step that packed the forward operands only
            self.repack(2)
        main = torch.cuda.current_stream(self.device)
        if self._wg_stream is None:
            self._wg_stream = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
            self.wg_ws2 = torch.empty_like(self.wg_ws)
        n_side = int(os.environ.get("ADN_WGRAD_STREAMS", "2")) if self.overlap_wgrad else 0
        sides = [main] if n_side == 0 else list(self._wg_stream[:n_side])
        wss = [self.wg_ws, self.wg_ws2]
        turn = [0]

        def next_side():                       # weight gradients alternate between the side streams, each with its own split-K workspace
            k = turn[0] % len(sides)
            turn[0] += 1
            return sides[k], wss[k].data_ptr()

        def join_sides():
            for sd in sides:
                main.wait_stream(sd)

        keep = []                              # tensors the side streams still read: alive until they have been joined

        def conv_bwd(layer, dy_ptr, dy_ld, need_dx=True):
            p, ci_, bi, l, c0, c1, co = layer
            hh, ww = hs[l], wz[l]
            pixels = n * hh * ww
            z, y, src0, src1 = sv[(p, ci_)]
            scale, shift, mean, invstd = self.stats[(p, ci_)]
            bn = f"{p}.double_conv.{bi}"
            dz = torch.empty_like(z)
            _lib.check(lib.adn_bn_relu_backward_bf16(dy_ptr, dy_ld, z.data_ptr(), pixels, co, scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                                     invstd.data_ptr(), self._gptr(f"{bn}.weight"), self._gptr(f"{bn}.bias"), dz.data_ptr(), wsp, s),
                       f"bn bwd {bn}")
            self.launch_count += 3
            wkey = f"{p}.double_conv.{ci_}.weight"
            if c0 + c1 == 1:
                _lib.check(lib.adn_conv3x3_c1_wgrad_f32(dz.data_ptr(), sv["x"].data_ptr(), n, hh, ww, self._gptr(wkey), wsp, s), "c1 wgrad")
                self.launch_count += 2
                return None
            # The weight gradient only needs dz and the saved input; nothing downstream reads it before the optimizer.  It runs on a
            # SIDE STREAM (a parallel branch of the captured graph), beside the data-gradient chain of the following layers: the deep
            # layers' kernels fill a fraction of the SMs each, so the two branches overlap instead of queueing.
            keep.append(dz)
            side, wg = next_side()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                _lib.check(lib.adn_conv3x3_wgrad_f32(dz.data_ptr(), co, src0.data_ptr(), c0, hh, ww, n, hh, ww, self._gptr(wkey), 0, c0 + c1, wg, side.cuda_stream), f"wgrad {wkey}")
            self.launch_count += 3
            if c1:
                side, wg = next_side()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    _lib.check(lib.adn_conv3x3_wgrad_f32(dz.data_ptr(), co, src1.data_ptr(), c1, hh, ww, n, hh, ww, self._gptr(wkey), c0, c0 + c1, wg, side.cuda_stream), f"wgrad {wkey}")
                self.launch_count += 3
            if not need_dx:
                return None
            dx = torch.empty((n, hh, ww, c0 + c1), **bf)
            _lib.check(lib.adn_conv3x3_affine_bf16(dz.data_ptr(), co, 0, 0, 0, 0, n, hh, ww, self.packed[(p, ci_)][1].data_ptr(), c0 + c1,
                                                   self.ones.data_ptr(), self.zeros.data_ptr(), 0, dx.data_ptr(), s), f"dgrad {wkey}")
            self.launch_count += 1
            return dx

        with torch.cuda.device(self.device):
            layers = list(self.layers)
            head_in = sv["head_in"]
            dy = torch.empty_like(head_in)
            db = torch.empty(64, dtype=torch.float32, device=self.device)          # element 0 = d out.bias
            _lib.check(lib.adn_head1x1_backward(head_in.data_ptr(), d_out.data_ptr(), self._pptr("out.weight"), n * hs[0] * wz[0], dy.data_ptr(),
                                                self._gptr("out.weight"), db.data_ptr(), wsp, s), "head bwd")
            _lib.check(lib.adn_copy_bytes(self._gptr("out.bias"), db.data_ptr(), 4, s), "out.bias grad")
            self.launch_count += 5
            d_skip = {}
            cur_dy, cur_ld = dy, 64
            # decoder, reverse order: upconv4 (level 0) ... upconv1 (level 3); self.layers[10:] are the decoder convs in forward order
            for i, l in zip((3, 2, 1, 0), (0, 1, 2, 3)):
                conv3 = layers[10 + 2 * i + 1]
                conv0 = layers[10 + 2 * i]
                c = _CH[l]
                d_a = conv_bwd(conv3, cur_dy.data_ptr(), cur_ld)
                d_cat = conv_bwd(conv0, d_a.data_ptr(), c)                       # (n, h, w, 2c): [skip | up]
                d_skip[l] = d_cat
                name = f"upconv{i + 1}.up"
                src, _up = sv[name]
                ci = _CH[l + 1]
                up_ptr = d_cat.data_ptr() + 2 * c
                _lib.check(lib.adn_channel_sum_f32(up_ptr, 2 * c, n * hs[l] * wz[l], c, self._gptr(f"{name}.bias"), wsp, s), "convT bias grad")
                side, wg = next_side()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    _lib.check(lib.adn_convt2x2_wgrad_f32(src.data_ptr(), ci, d_cat.data_ptr(), 2 * c, c, c, n, hs[l + 1], wz[l + 1],
                                                          self._gptr(f"{name}.weight"), wg, side.cuda_stream), "convT wgrad")
                d_src = torch.empty((n, hs[l + 1], wz[l + 1], ci), **bf)
                _lib.check(lib.adn_convt2x2_dgrad_bf16(d_cat.data_ptr(), 2 * c, c, c, n, hs[l + 1], wz[l + 1], self.packed[name][1].data_ptr(), ci,
                                                       d_src.data_ptr(), s), "convT dgrad")
                self.launch_count += 4
                cur_dy, cur_ld = d_src, ci
            # DDP: the decoder's gradients (flat slice from upconv1.up.weight to the end, 39 % of the parameters) are final: their
            # all-reduce runs under the bottleneck / encoder backward
            join_sides()
            self._start_bucket(self.offsets["upconv1.up.weight"][0], self.numel)
            # bottleneck
            d_a = conv_bwd(layers[9], cur_dy.data_ptr(), cur_ld)
            d_pool = conv_bwd(layers[8], d_a.data_ptr(), 1024)
            join_sides()
            self._start_bucket(self.offsets["bottleneck.double_conv.0.weight"][0], self.offsets["upconv1.up.weight"][0])   # 46 %
            # encoder, levels 3..0
            for l in (3, 2, 1, 0):
                c = _CH[l]
                _z, y_skip, _s0, _s1 = sv[layers[2 * l + 1][0], 3]
                dy_s = torch.empty_like(y_skip)
                _lib.check(lib.adn_maxpool2x2_backward_add_bf16(y_skip.data_ptr(), d_pool.data_ptr(), d_skip[l].data_ptr(), 2 * c, n, hs[l], wz[l], c,
                                                                dy_s.data_ptr(), s), "maxpool bwd")
                self.launch_count += 1
                d_a = conv_bwd(layers[2 * l + 1], dy_s.data_ptr(), c)
                d_pool = conv_bwd(layers[2 * l], d_a.data_ptr(), c, need_dx=(l > 0))
            join_sides()
            keep.clear()
            self._start_bucket(0, self.offsets["bottleneck.double_conv.0.weight"][0])                                    # 15 %
        self.saved = None

    # ------------------------------------------------------------------ loss (loss.py:83-95)
    def loss_and_grad(self, pred, target):
        """-> (4 loss values on the device: total, stft, mel, l1;  d total / d pred)."""
        lib, s = self.lib, _lib.stream_ptr()
        b, _, f, t = pred.shape
        target = target.float().contiguous()
        need = max(int(lib.adn_loss_workspace_bytes(b, f, t)), int(lib.adn_loss_backward_workspace_bytes(b, f, t)))
        if self._loss_ws is None or self._loss_ws.numel() < need:
            self._loss_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        d_pred = torch.empty_like(pred)
        with torch.cuda.device(self.device):
            _lib.check(lib.adn_combined_loss_f32(pred.data_ptr(), target.data_ptr(), b, f, t, self.mel_fb.data_ptr(), self._loss_ws.data_ptr(),
                                                 self.loss_out.data_ptr(), s), "loss fwd")
            _lib.check(lib.adn_combined_loss_backward_f32(pred.data_ptr(), target.data_ptr(), b, f, t, self.mel_fb.data_ptr(), 0.4, 0.4, 0.2,
                                                          self._loss_ws.data_ptr(), d_pred.data_ptr(), s), "loss bwd")
        self.launch_count += 6
        losses = torch.empty_like(self.loss_out)
        with torch.cuda.device(self.device):
            _lib.check(lib.adn_copy_bytes(losses.data_ptr(), self.loss_out.data_ptr(), 16, s), "loss copy")
        return losses, d_pred

    # ------------------------------------------------------------------ optimizer (train.py:66,70,71)
    def zero_grad(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.adn_zero_bytes(self.G.data_ptr(), 4 * self.numel, _lib.stream_ptr()), "zero_grad")

    def _start_bucket(self, lo: int, hi: int):
        """DDP: start the asynchronous all-reduce(AVG) of G[lo:hi] (three buckets per step, launched from backward() in the order
        the gradients become final: decoder, bottleneck, encoder).  No-op without an initialised process group or in an
        autograd-driven backward (torch's own DDP wrapper / the reference loop average there)."""
        if not self._bucketed:
            return
        from .sharding import average_gradients_async
        self._buckets.append(average_gradients_async(self.G[lo:hi], self.group))

    def all_reduce_grads(self):
        """DDP gradient averaging (31.04 M fp32 = 124 MB): wait for the buckets backward() launched, or -- when the step did not go
        through train_step -- one all-reduce over the whole flat buffer."""
        if self._buckets:
            for b in self._buckets:
                b.wait()
            self._buckets = []
            return
        from .sharding import average_gradients_
        average_gradients_(self.G, self.group)

    def optimizer_step(self, split_pack: bool = False):
        """clip_grad_norm_(max_norm) + AdamW.step(), then refresh the bf16 operand copies.  Returns the gradient norm (device scalar)."""
        lib, s = self.lib, _lib.stream_ptr()
        self.step_count += 1
        with torch.cuda.device(self.device):
            _lib.check(lib.adn_grad_norm_f32(self.G.data_ptr(), self.numel, self.max_norm, self.norm_coef.data_ptr(), self.ws.data_ptr(), s), "grad norm")
            _lib.check(lib.adn_adamw_step_dev_f32(self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(), self.V.data_ptr(), self.numel,
                                                  self.norm_coef.data_ptr(), self.step_dev.data_ptr(), self.lr, self.betas[0], self.betas[1],
                                                  self.eps, self.weight_decay, s), "adamw")
        self.launch_count += 4
        self.repack(1 if split_pack else 3)
        return self.norm_coef[0]

    def train_step(self, noisy, clean):
        """One iteration of train_one_epoch (train.py:65-72).  Returns the device tensor (total, stft, mel, l1)."""
        # zeroing the gradient buffer and the data-gradient weight packs (stale since the last step's AdamW) are first needed in
        # backward: they run on a side stream under the forward
        main = torch.cuda.current_stream(self.device)
        if self._aux_stream is None:
            self._aux_stream = torch.cuda.Stream(self.device)
        aux = self._aux_stream if self.split_pack else main
        aux.wait_stream(main)
        with torch.cuda.stream(aux):
            self.zero_grad()
            if self._dgrad_stale:
                self.repack(2)
        out = self.forward(noisy)
        losses, d_pred = self.loss_and_grad(out, clean)
        main.wait_stream(aux)
        self._bucketed = True
        try:
            self.backward(d_pred)
        finally:
            self._bucketed = False
        self.all_reduce_grads()
        self.optimizer_step(split_pack=self.split_pack)
        return losses

    def release_graph(self):
        """Drop the captured step.  A CUDA graph that contains NCCL kernels must be destroyed BEFORE its process group is
        (destroy_process_group with such a graph alive deadlocks); call this before tearing the group down."""
        self._graph = None

    def train_step_graphed(self, noisy, clean):
        """train_step through ONE CUDA graph: the step is ~330 small launches, so replaying a captured graph removes the host
        launch cost.  The first call for a shape runs eagerly (it also performs the one-off cudaFuncSetAttribute calls, which
        cannot be captured), the second captures and replays, later calls copy the batch into the static inputs and replay."""
        key = (tuple(noisy.shape), noisy.device)
        g = self._graph
        if g is None or g["key"] != key:
            if g is None or g.get("warm") != key:
                self._graph = {"key": None, "warm": key}
                return self.train_step(noisy, clean)
            static_in = (torch.empty_like(noisy, dtype=torch.float32), torch.empty_like(clean, dtype=torch.float32))
            static_in[0].copy_(noisy); static_in[1].copy_(clean)
            graph = torch.cuda.CUDAGraph()
            count0, launches0 = self.step_count, self.launch_count
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                losses = self.train_step(*static_in)
            self.step_count = count0                       # capture records the step, it does not execute it
            self._graph = g = {"key": key, "warm": key, "graph": graph, "in": static_in, "out": losses,
                               "launches": self.launch_count - launches0}
            self.launch_count = launches0
        else:
            g["in"][0].copy_(noisy, non_blocking=True); g["in"][1].copy_(clean, non_blocking=True)
        g["graph"].replay()
        self.step_count += 1
        self.launch_count += g["launches"]
        # The replay updated the parameters and the BatchNorm running statistics through raw device pointers: no tensor version
        # counter moved, so the module's eval-mode cache (packed bf16 weights + folded BN scale / shift, keyed on versions) must be
        # dropped by hand, or model.eval() after graphed steps would run with the weights of the last eager step.
        self.model._packed = None
        return g["out"]

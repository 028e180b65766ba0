"""Drop-in for the reference's code/model.py ``UNet`` -- same constructor, same 136-entry ``state_dict`` (so existing
``unet_denoiser_{noise}.pth`` checkpoints load, test.py:63-66), same ``forward(x)`` contract ((N,1,F,T) float32 in,
(N,1,F,T) float32 out on the input's device) -- whose forward runs the hand-written sm_100a kernels of
libadn_b200.so: bf16 operands, fp32 accumulation in TMEM, fp32 BatchNorm/ReLU epilogues.

The module holds fp32 master parameters in the reference layout; a packed copy (bf16 K-major weights, folded BN
scale/shift) is rebuilt whenever they change.  In ``train()`` mode ``forward`` runs the batch-statistics kernels of
``training.TrainEngine`` and returns a tensor that is part of the autograd graph: ``loss.backward()`` runs the hand-written
backward kernels and deposits ``.grad`` on every parameter, so the reference's train_one_epoch body (train.py:65-72) works
unchanged with any torch optimizer.  No CPU code path: a host tensor (test.py:100,113) is copied to the GPU, processed by the
same kernels and copied back; without a CUDA device every call raises.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib
from .checkpoint import BN_EPS, state_dict_spec

_MAX_CHUNK = 64          # images per kernel batch (bounds the activation workspace; test.py:113 feeds the whole set)


class _Holder(nn.Module):
    """Attribute container: lets parameters sit under the reference's dotted names."""


def _attach(root: nn.Module, dotted: str, tensor: torch.Tensor, is_buffer: bool) -> None:
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, _Holder())
        mod = getattr(mod, p)
    if is_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor))


class _TrainForward(torch.autograd.Function):
    """model(x) in train() mode as one autograd node: forward = TrainEngine.forward, backward = TrainEngine.backward."""

    @staticmethod
    def forward(ctx, x, engine, *params):
        ctx.engine = engine
        ctx.n_params = len(params)
        return engine.forward(x)

    @staticmethod
    def backward(ctx, d_out):
        eng = ctx.engine
        eng.zero_grad()
        eng.backward(d_out)
        grads = tuple(eng.gview(k).clone() for k in eng.offsets)      # autograd accumulates these into param.grad
        return (None, None) + grads


class UNet(nn.Module):
    """model.py:53-94.  ``UNet(in_channels=1, num_classes=1)``."""

    def __init__(self, in_channels=1, num_classes=1):
        super().__init__()
        if in_channels != 1 or num_classes != 1:
            raise ValueError("the B200 build is specialised to the reference's UNet(in_channels=1, num_classes=1)")
        # Same initialisation as the reference's nn modules, drawn from the GLOBAL torch RNG in the same order and with the same
        # calls (nn.Conv2d / nn.ConvTranspose2d.reset_parameters: kaiming_uniform_(a=sqrt(5)) weights, U(+-1/sqrt(fan_in)) biases;
        # BatchNorm2d ones / zeros): torch.manual_seed(s); UNet() yields the state_dict the reference's UNet() would.
        import math
        last_fan_in = 1
        for key, (shape, dtype) in state_dict_spec().items():
            leaf = key.rsplit(".", 1)[1]
            is_buffer = leaf in ("running_mean", "running_var", "num_batches_tracked")
            is_bn = ".double_conv.1." in key or ".double_conv.4." in key
            if dtype == torch.int64:
                t = torch.zeros((), dtype=torch.int64)
            elif is_bn:
                t = torch.ones(shape) if leaf in ("weight", "running_var") else torch.zeros(shape)
            elif len(shape) == 4:
                t = torch.empty(shape)
                nn.init.kaiming_uniform_(t, a=math.sqrt(5))
                last_fan_in = shape[1] * shape[2] * shape[3]
            else:                                                  # the bias that follows its conv / conv-transpose / head weight
                bound = 1.0 / math.sqrt(last_fan_in)
                t = torch.empty(shape)
                nn.init.uniform_(t, -bound, bound)
            _attach(self, key, t, is_buffer)
        self._packed = None
        self._packed_key = None
        self._engine = None
        self._ws = {}
        # ConvTranspose2d + first decoder conv as one merged-weight kernel where the level has >= 128 output channels (ADN_UPMERGE=0: off)
        self.upmerge = os.environ.get("ADN_UPMERGE", "1") != "0"
        # the 128-output-channel convs without a fused pool in the two-classes-per-tile parity formulation (ADN_PLANE128=0: off)
        self.plane128 = os.environ.get("ADN_PLANE128", "1") != "0"
        self.profile = None      # set to a list to record (layer, kind, flops, start_event, end_event) per kernel call
        self.launch_count = 0    # kernels launched by this module so far (bench.py's gpu_launches)

    # ------------------------------------------------------------------ packing
    def _version_key(self, device):
        return (str(device),) + tuple(int(v._version) for v in self.state_dict(keep_vars=True).values())

    def _pack(self, device):
        """fp32 reference-layout tensors -> bf16 [Co][tap][Ci] weights + folded fp32 BN scale/shift, on `device`."""
        lib = _lib.load()
        s = _lib.stream_ptr()
        sd = {k: v.detach().to(device=device) for k, v in self.state_dict(keep_vars=True).items()}
        packed = {}

        def conv(prefix, idx_conv, idx_bn, first=False):
            w = sd[f"{prefix}.double_conv.{idx_conv}.weight"].float().contiguous()
            co, ci = w.shape[0], w.shape[1]
            scale = torch.empty(co, dtype=torch.float32, device=device)
            shift = torch.empty(co, dtype=torch.float32, device=device)
            bn = f"{prefix}.double_conv.{idx_bn}"
            args = [sd[f"{prefix}.double_conv.{idx_conv}.bias"], sd[f"{bn}.weight"], sd[f"{bn}.bias"], sd[f"{bn}.running_mean"],
                    sd[f"{bn}.running_var"]]
            args = [a.float().contiguous() for a in args]
            _lib.check(lib.adn_fold_bn_f32(*[a.data_ptr() for a in args], BN_EPS, co, scale.data_ptr(), shift.data_ptr(), s), "adn_fold_bn_f32")
            if first:
                return {"w": w, "scale": scale, "shift": shift, "co": co, "ci": ci, "keep": args}
            wp = torch.empty((co, 9, ci), dtype=torch.bfloat16, device=device)
            _lib.check(lib.adn_pack_conv3x3_weight_bf16(w.data_ptr(), co, ci, wp.data_ptr(), s), "adn_pack_conv3x3_weight_bf16")
            d = {"w": wp, "scale": scale, "shift": shift, "co": co, "ci": ci, "keep": (w, args)}
            if co == 128 and ci >= 128:         # parity-class formulation with two classes per tile (conv3x3_upm2_kernel without a low tensor)
                d["bsh"] = torch.empty(int(lib.adn_upmerged_pair_weight_elems(128, ci, 0, 0)), dtype=torch.bfloat16, device=device)
                d["b1"] = torch.empty(int(lib.adn_upmerged_pair_weight_elems(128, ci, 0, 1)), dtype=torch.bfloat16, device=device)
                _lib.check(lib.adn_pack_upmerged_pair_weight_bf16(wp.data_ptr(), 128, ci, 0, d["bsh"].data_ptr(), d["b1"].data_ptr(), s),
                           "adn_pack_upmerged_pair_weight_bf16")
            return d

        for name in ("downconv1", "downconv2", "downconv3", "downconv4"):
            packed[f"{name}.0"] = conv(f"{name}.conv", 0, 1, first=(name == "downconv1"))
            packed[f"{name}.3"] = conv(f"{name}.conv", 3, 4)
        packed["bottleneck.0"] = conv("bottleneck", 0, 1)
        packed["bottleneck.3"] = conv("bottleneck", 3, 4)
        for name in ("upconv1", "upconv2", "upconv3", "upconv4"):
            w = sd[f"{name}.up.weight"].float().contiguous()
            ci, co = w.shape[0], w.shape[1]
            wp = torch.empty((4, co, ci), dtype=torch.bfloat16, device=device)
            _lib.check(lib.adn_pack_convt2x2_weight_bf16(w.data_ptr(), ci, co, wp.data_ptr(), s), "adn_pack_convt2x2_weight_bf16")
            packed[f"{name}.up"] = {"w": wp, "bias": sd[f"{name}.up.bias"].float().contiguous(), "co": co, "ci": ci, "keep": w}
            packed[f"{name}.0"] = conv(f"{name}.conv", 0, 1)
            packed[f"{name}.3"] = conv(f"{name}.conv", 3, 4)
            # ConvTranspose2d merged into the first conv's weights (csrc/conv_upm.cu) for the levels with >= 128 output channels.
            # Measured at batch 64 x (257,1034): upconv1 0.26 + 1.64 -> 1.63 ms, upconv2 0.32 + 1.58 -> 1.51 ms.  The 128-wide level
            # LOSES with one parity class per tile (0.38 + 1.65 -> 2.36 ms: every class re-reads all four skip planes, and a
            # 256x128x16 UMMA is too short to hide 78 KB of TMA writes per chunk), so it runs the two-classes-per-tile kernel.
            pc = packed[f"{name}.0"]
            if pc["co"] % 128 == 0:
                w3, bt = pc["keep"][0], packed[f"{name}.up"]["bias"]
                c0 = pc["ci"] - co
                wm = torch.empty((pc["co"], 9 * c0 + 16 * ci), dtype=torch.bfloat16, device=device)
                shift_m = torch.empty(pc["co"], dtype=torch.float32, device=device)
                wb = torch.empty((9, pc["co"]), dtype=torch.float32, device=device)
                _lib.check(lib.adn_pack_upmerged_weight_bf16(w3.data_ptr(), w.data_ptr(), bt.data_ptr(), pc["scale"].data_ptr(),
                                                             pc["shift"].data_ptr(), pc["co"], c0, co, ci, wm.data_ptr(),
                                                             shift_m.data_ptr(), wb.data_ptr(), s), "adn_pack_upmerged_weight_bf16")
                pm = {"w": wm, "scale": pc["scale"], "shift": shift_m, "wb": wb, "co": pc["co"], "c0": c0, "cl": ci}
                if pc["co"] == 128:
                    bsh = torch.empty(int(lib.adn_upmerged_pair_weight_elems(128, c0, ci, 0)), dtype=torch.bfloat16, device=device)
                    b1 = torch.empty(int(lib.adn_upmerged_pair_weight_elems(128, c0, ci, 1)), dtype=torch.bfloat16, device=device)
                    _lib.check(lib.adn_pack_upmerged_pair_weight_bf16(wm.data_ptr(), 128, c0, ci, bsh.data_ptr(), b1.data_ptr(), s),
                               "adn_pack_upmerged_pair_weight_bf16")
                    pm["bsh"], pm["b1"] = bsh, b1
                if pc["co"] % 256 == 0 or pc["co"] == 128:
                    packed[f"{name}.m"] = pm
        packed["out"] = {"w": sd["out.weight"].float().reshape(-1).contiguous(), "b": sd["out.bias"].float().contiguous()}
        return packed

    def packed(self, device):
        key = self._version_key(device)
        if self._packed is None or self._packed_key != key:
            with torch.cuda.device(device):
                self._packed = self._pack(device)
            self._packed_key = key
            self._ws = {}
        return self._packed

    # ------------------------------------------------------------------ forward
    def _workspace(self, n, h, w, device):
        key = (n, h, w, str(device))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        bf = dict(dtype=torch.bfloat16, device=device)
        hs, wsz = [h], [w]
        for _ in range(4):
            hs.append(hs[-1] // 2)
            wsz.append(wsz[-1] // 2)
        if hs[4] < 1 or wsz[4] < 1:
            raise ValueError(f"input {h}x{w} is too small for four 2x2 poolings")
        ch = [64, 128, 256, 512, 1024]
        ws = {"hs": hs, "ws": wsz}
        for l in range(5):
            ws[f"a{l}"] = torch.empty((n, hs[l], wsz[l], ch[l]), **bf)      # first conv of the level
            ws[f"s{l}"] = torch.empty((n, hs[l], wsz[l], ch[l]), **bf)      # second conv = skip / bottleneck out
            if l < 4:
                ws[f"p{l}"] = torch.empty((n, hs[l + 1], wsz[l + 1], ch[l]), **bf)          # pooled
                ws[f"ua{l}"] = torch.empty((n, hs[l], wsz[l], ch[l]), **bf)
                if l > 0:                      # level 0's second decoder conv feeds the fused 1x1 head instead
                    ws[f"ub{l}"] = torch.empty((n, hs[l], wsz[l], ch[l]), **bf)
        self._ws = {key: ws}          # keep one shape resident
        return ws

    def _forward_chunk(self, x, out, pk):
        lib = _lib.load()
        s = _lib.stream_ptr()
        n, _, h, w = x.shape
        ws = self._workspace(n, h, w, x.device)
        hs, wz = ws["hs"], ws["ws"]

        prof = self.profile

        def timed(layer, kind, flops, launches, fn, *args):
            if prof is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            st = fn(*args)
            if prof is not None:
                e1.record()
                prof.append((layer, kind, flops, e0, e1))
            self.launch_count += launches
            _lib.check(st, f"{fn.__name__}[{layer}]")

        def conv(layer, src0, c0, src1, c1, h1, w1, lvl, dst, pool=None):
            p = pk[layer]
            flops = 2.0 * n * hs[lvl] * wz[lvl] * p["co"] * 9 * (c0 + c1)
            # measured at batch 64 x (257,1034): upconv3.3 (128 -> 128) 0.89 -> 0.80 ms; downconv2.0 (64 -> 128, one chunk) 0.61 -> 0.62: not used
            if self.plane128 and "bsh" in p and src1 is None and c0 >= 128 and hs[lvl] >= 2 and wz[lvl] >= 2:
                if pool is None:
                    timed(layer, "conv3x3", flops, 1, lib.adn_conv3x3_upmerged_pair_bn_relu_bf16, src0.data_ptr(), c0, 0, 0, 0, 0, n, hs[lvl],
                          wz[lvl], p["bsh"].data_ptr(), p["b1"].data_ptr(), 128, p["scale"].data_ptr(), p["shift"].data_ptr(), 0,
                          dst.data_ptr(), s)
                else:                  # fused pool: running maximum over the four parity classes of a half-resolution pixel
                    timed(layer, "conv3x3", flops, 1, lib.adn_conv3x3_pair_bn_relu_pool_bf16, src0.data_ptr(), c0, n, hs[lvl], wz[lvl],
                          p["bsh"].data_ptr(), p["b1"].data_ptr(), 128, p["scale"].data_ptr(), p["shift"].data_ptr(), dst.data_ptr(),
                          pool.data_ptr(), s)
                return
            # the 2x2 max-pool of DownSampleLayer (model.py:31) is fused into the conv epilogue when `pool` is given
            timed(layer, "conv3x3", flops, 1, lib.adn_conv3x3_bn_relu_bf16, src0.data_ptr(), c0,
                  src1.data_ptr() if src1 is not None else 0, c1, h1, w1, n, hs[lvl], wz[lvl], p["w"].data_ptr(), p["co"],
                  p["scale"].data_ptr(), p["shift"].data_ptr(), dst.data_ptr(), pool.data_ptr() if pool is not None else 0, s)

        # encoder (DownSampleLayer.forward, model.py:29-32)
        p = pk["downconv1.0"]
        timed("downconv1.0", "conv3x3_c1", 2.0 * n * h * w * 64 * 9, 1, lib.adn_conv3x3_c1_bn_relu_bf16, x.data_ptr(), n, h, w,
              p["w"].data_ptr(), p["scale"].data_ptr(), p["shift"].data_ptr(), ws["a0"].data_ptr(), s)
        conv("downconv1.3", ws["a0"], 64, None, 0, 0, 0, 0, ws["s0"], ws["p0"])
        ch = [64, 128, 256, 512, 1024]
        for l in (1, 2, 3):
            conv(f"downconv{l + 1}.0", ws[f"p{l - 1}"], ch[l - 1], None, 0, 0, 0, l, ws[f"a{l}"])
            conv(f"downconv{l + 1}.3", ws[f"a{l}"], ch[l], None, 0, 0, 0, l, ws[f"s{l}"], ws[f"p{l}"])
        conv("bottleneck.0", ws["p3"], 512, None, 0, 0, 0, 4, ws["a4"])
        conv("bottleneck.3", ws["a4"], 1024, None, 0, 0, 0, 4, ws["s4"])
        # decoder (UpSampleLayer.forward, model.py:41-50)
        cur = ws["s4"]
        for i, l in enumerate((3, 2, 1, 0)):
            name = f"upconv{i + 1}"
            pu = pk[f"{name}.up"]
            pm = pk.get(f"{name}.m") if self.upmerge else None
            if pm is not None:
                # executed MACs: 9 taps over the skip channels + 4 merged taps over the low-resolution channels
                flops = 2.0 * n * hs[l] * wz[l] * pm["co"] * (9 * pm["c0"] + 4 * pm["cl"])
                if "bsh" in pm:
                    timed(f"{name}.up+0", "conv3x3_upm", flops, 1, lib.adn_conv3x3_upmerged_pair_bn_relu_bf16, ws[f"s{l}"].data_ptr(),
                          pm["c0"], cur.data_ptr(), pm["cl"], hs[l + 1], wz[l + 1], n, hs[l], wz[l], pm["bsh"].data_ptr(), pm["b1"].data_ptr(),
                          pm["co"], pm["scale"].data_ptr(), pm["shift"].data_ptr(), pm["wb"].data_ptr(), ws[f"ua{l}"].data_ptr(), s)
                else:
                    timed(f"{name}.up+0", "conv3x3_upm", flops, 1, lib.adn_conv3x3_upmerged_bn_relu_bf16, ws[f"s{l}"].data_ptr(), pm["c0"],
                          cur.data_ptr(), pm["cl"], hs[l + 1], wz[l + 1], n, hs[l], wz[l], pm["w"].data_ptr(), pm["co"],
                          pm["scale"].data_ptr(), pm["shift"].data_ptr(), pm["wb"].data_ptr(), ws[f"ua{l}"].data_ptr(), s)
                conv(f"{name}.3", ws[f"ua{l}"], ch[l], None, 0, 0, 0, l, ws[f"ub{l}"])
                cur = ws[f"ub{l}"]
                continue
            if f"u{l}" not in ws:            # the up-sampled map exists only where the ConvTranspose is not merged into the conv (level 0)
                ws[f"u{l}"] = torch.empty((n, 2 * hs[l + 1], 2 * wz[l + 1], ch[l]), dtype=torch.bfloat16, device=x.device)
            timed(f"{name}.up", "convt2x2", 2.0 * n * hs[l + 1] * wz[l + 1] * pu["ci"] * 4 * pu["co"], 1, lib.adn_convt2x2_bf16,
                  cur.data_ptr(), pu["ci"], n, hs[l + 1], wz[l + 1], pu["w"].data_ptr(), pu["co"], pu["bias"].data_ptr(),
                  ws[f"u{l}"].data_ptr(), s)
            conv(f"{name}.0", ws[f"s{l}"], ch[l], ws[f"u{l}"], ch[l], 2 * hs[l + 1], 2 * wz[l + 1], l, ws[f"ua{l}"])
            if l > 0:
                conv(f"{name}.3", ws[f"ua{l}"], ch[l], None, 0, 0, 0, l, ws[f"ub{l}"])
                cur = ws[f"ub{l}"]
            else:
                p = pk[f"{name}.3"]
                timed(f"{name}.3+out", "conv3x3_head", 2.0 * n * h * w * 64 * (9 * 64 + 1), 1, lib.adn_conv3x3_bn_relu_head_f32,
                      ws["ua0"].data_ptr(), 64, 0, 0, 0, 0, n, h, w, p["w"].data_ptr(), 64, p["scale"].data_ptr(),
                      p["shift"].data_ptr(), pk["out"]["w"].data_ptr(), pk["out"]["b"].data_ptr(), out.data_ptr(), s)

    def train_engine(self, device=None, **optimizer_kwargs):
        """The TrainEngine bound to this module (created on first use; the module's parameters then alias its flat buffer)."""
        from .training import TrainEngine
        if self._engine is None or (device is not None and torch.device(device) != self._engine.device):
            self._engine = TrainEngine(self, device=device, **optimizer_kwargs)
        elif optimizer_kwargs:
            self._engine.set_hyperparameters(**optimizer_kwargs)
        return self._engine

    def _compute_device(self):
        """Where a HOST input is processed: the parameters' device when the module was moved to a GPU, else the current one."""
        p = next(self.parameters())
        return p.device if p.is_cuda else torch.device("cuda", torch.cuda.current_device())

    def forward(self, x):
        """UNet.forward, model.py:70-94: eval() -> running-statistics BatchNorm folded into the conv epilogues;
        train() -> batch-statistics BatchNorm, differentiable (train.py:67,69).

        The result lives on the input's device, as with the reference module.  A HOST tensor -- what test.py:100,113 feeds
        (`model` is never moved off the CPU there) -- is staged through pinned memory to the GPU, run through the same
        kernels, and the result is copied back to a host tensor: host in / host out is not a CPU code path."""
        _lib.require_cuda()
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError("expected (N, 1, F, T) input")
        if not x.is_cuda:
            dev = self._compute_device()
            y = self.forward(_host_to_device(x, dev))
            return y.cpu()
        if self.training:
            eng = self.train_engine(x.device)
            eng.refresh_if_parameters_changed()
            if torch.is_grad_enabled():
                return _TrainForward.apply(x, eng, *[p for _, p in self.named_parameters()])
            return eng.forward(x)
        x = x.float().contiguous()
        with torch.cuda.device(x.device), torch.no_grad():
            pk = self.packed(x.device)
            out = torch.empty_like(x)
            for i in range(0, x.shape[0], _MAX_CHUNK):
                self._forward_chunk(x[i:i + _MAX_CHUNK], out[i:i + _MAX_CHUNK], pk)
        return out


def _host_to_device(x: torch.Tensor, dev) -> torch.Tensor:
    """Host tensor -> device, through a pinned staging copy when the source is pageable and large (a pageable cudaMemcpy is
    staged by the driver in small chunks); keeps autograd history (``.to`` is differentiable)."""
    if x.requires_grad or x.numel() < (1 << 16) or x.is_pinned():
        return x.to(dev, non_blocking=x.is_pinned())
    return x.float().contiguous().pin_memory().to(dev, non_blocking=True)

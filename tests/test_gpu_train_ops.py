"""Parity of the training-step kernels (through the C ABI) against torch CPU fp32 autograd on identical bf16-rounded operands.

Each test restates ONE op of the reference's train.py:65-72 body (model.py layers in train() mode, loss.py backward,
clip_grad_norm_, AdamW) with plain torch on the CPU and compares.  Tolerances: operands are bf16, accumulation fp32, so a
result that is itself stored in bf16 is within 2^-8 relative of the fp32 answer; fp32 results (weight gradients, statistics)
are held to 2e-3 of the tensor's scale."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from audiodenoiser_b200 import _lib
from audiodenoiser_b200.loss import mel_filterbank
from oracle import loss_oracle

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def nhwc(x):
    """(N,C,H,W) fp32 cpu -> (N,H,W,C) bf16 cuda"""
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev())


def nchw(x):
    """(N,H,W,C) bf16 cuda -> (N,C,H,W) fp32 cpu"""
    return x.float().cpu().permute(0, 3, 1, 2).contiguous()


def bf(x):
    return x.to(torch.bfloat16).float()


def rel(got, ref):
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-30))


def ws():
    lib = _lib.load()
    return torch.empty(int(lib.adn_train_workspace_bytes()), dtype=torch.uint8, device=dev())


def sp():
    return _lib.stream_ptr()


def wgws():
    return torch.empty(int(_lib.load().adn_wgrad_workspace_bytes()), dtype=torch.uint8, device=dev())


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 16, 16, 64, 64), (1, 32, 8, 128, 64), (2, 8, 24, 64, 256)])
def test_conv3x3_affine_no_relu(n, h, w, ci, co):
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    x = bf(torch.randn(n, ci, h, w, generator=g))
    wt = bf(torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (9 * ci)) ** 0.5)
    bias = torch.randn(co, generator=g)
    ref = F.conv2d(x, wt, bias, padding=1)
    xd = nhwc(x)
    wp = torch.empty((co, 9, ci), dtype=torch.bfloat16, device=dev())
    wd = wt.to(dev())
    _lib.check(lib.adn_pack_conv3x3_weight_bf16(wd.data_ptr(), co, ci, wp.data_ptr(), sp()))
    ones = torch.ones(co, device=dev()); b = bias.to(dev())
    out = torch.empty((n, h, w, co), dtype=torch.bfloat16, device=dev())
    _lib.check(lib.adn_conv3x3_affine_bf16(xd.data_ptr(), ci, 0, 0, 0, 0, n, h, w, wp.data_ptr(), co, ones.data_ptr(), b.data_ptr(), 0,
                                           out.data_ptr(), sp()))
    assert (nchw(out) < 0).any()                       # no ReLU
    assert rel(nchw(out), ref) < 6e-3


@pytest.mark.parametrize("pixels_shape,c", [((2, 16, 8), 64), ((3, 8, 8), 256), ((1, 4, 4), 1024)])
def test_bn_train_stats_apply_and_backward(pixels_shape, c):
    lib = _lib.load()
    n, h, w = pixels_shape
    g = torch.Generator().manual_seed(2)
    z = bf(torch.randn(n, c, h, w, generator=g) * 1.7 + 0.3)
    gamma = (0.75 + 0.5 * torch.rand(c, generator=g)).requires_grad_()
    beta = (0.1 * torch.randn(c, generator=g)).requires_grad_()
    rm0, rv0 = 0.1 * torch.randn(c, generator=g), 0.5 + torch.rand(c, generator=g)
    zr = z.clone().requires_grad_()
    rm, rv = rm0.clone(), rv0.clone()
    y_ref = F.relu(F.batch_norm(zr, rm, rv, gamma, beta, training=True, momentum=0.1, eps=1e-5))
    dy = bf(torch.randn(n, c, h, w, generator=g))
    y_ref.backward(dy)

    d = dev()
    zd = nhwc(z)
    pixels = n * h * w
    f32 = lambda: torch.empty(c, dtype=torch.float32, device=d)
    scale, shift, mean, invstd = f32(), f32(), f32(), f32()
    rmd, rvd = rm0.to(d), rv0.to(d)
    wsp = ws()
    gd, bd = gamma.detach().to(d), beta.detach().to(d)          # keep every device operand alive across the call
    _lib.check(lib.adn_bn_train_stats_f32(zd.data_ptr(), pixels, c, gd.data_ptr(), bd.data_ptr(), 1e-5, 0.1,
                                          rmd.data_ptr(), rvd.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                          invstd.data_ptr(), wsp.data_ptr(), sp()))
    assert torch.allclose(rmd.cpu(), rm, rtol=1e-4, atol=1e-5) and torch.allclose(rvd.cpu(), rv, rtol=1e-4, atol=1e-5)
    yd = torch.empty_like(zd)
    _lib.check(lib.adn_bn_relu_apply_bf16(zd.data_ptr(), scale.data_ptr(), shift.data_ptr(), pixels, c, yd.data_ptr(), sp()))
    assert rel(nchw(yd), y_ref.detach()) < 5e-3
    # backward; dy given as a channel slice of a wider tensor (pixel stride 2c)
    wide = torch.zeros((n, h, w, 2 * c), dtype=torch.bfloat16, device=d)
    wide[..., c:] = nhwc(dy)
    dz = torch.empty_like(zd)
    dg, db = f32(), f32()
    dy_ptr = wide.data_ptr() + 2 * c
    _lib.check(lib.adn_bn_relu_backward_bf16(dy_ptr, 2 * c, zd.data_ptr(), pixels, c, scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                             invstd.data_ptr(), dg.data_ptr(), db.data_ptr(), dz.data_ptr(), wsp.data_ptr(), sp()))
    assert rel(dg.cpu(), gamma.grad) < 5e-3 and rel(db.cpu(), beta.grad) < 5e-3
    assert rel(nchw(dz), zr.grad) < 8e-3


def test_maxpool_backward_add():
    lib = _lib.load()
    n, h, w, c = 2, 8, 12, 64
    g = torch.Generator().manual_seed(3)
    y = bf(F.relu(torch.randn(n, c, h, w, generator=g)))          # post-ReLU: plenty of ties at zero
    yr = y.clone().requires_grad_()
    dp = bf(torch.randn(n, c, h // 2, w // 2, generator=g))
    dd = bf(torch.randn(n, c, h, w, generator=g))
    F.max_pool2d(yr, 2).backward(dp)
    ref = yr.grad + dd
    out = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=dev())
    yd, dpd, ddd = nhwc(y), nhwc(dp), nhwc(dd)
    _lib.check(lib.adn_maxpool2x2_backward_add_bf16(yd.data_ptr(), dpd.data_ptr(), ddd.data_ptr(), c, n, h, w, c, out.data_ptr(), sp()))
    assert rel(nchw(out), bf(ref)) < 1e-6 or rel(nchw(out), ref) < 5e-3


def test_head_forward_backward():
    lib = _lib.load()
    n, h, w = 2, 16, 8
    g = torch.Generator().manual_seed(4)
    y = bf(torch.randn(n, 64, h, w, generator=g))
    wt = (torch.randn(1, 64, 1, 1, generator=g) * 0.2).requires_grad_()
    b = torch.randn(1, generator=g).requires_grad_()
    yr = y.clone().requires_grad_()
    out_ref = F.conv2d(yr, wt, b)
    d_out = torch.randn(n, 1, h, w, generator=g)
    out_ref.backward(d_out)
    d = dev()
    pixels = n * h * w
    yd = nhwc(y)
    wd, bd = wt.detach().reshape(-1).to(d), b.detach().to(d)
    out = torch.empty((n, 1, h, w), dtype=torch.float32, device=d)
    _lib.check(lib.adn_head1x1_forward_f32(yd.data_ptr(), wd.data_ptr(), bd.data_ptr(), pixels, out.data_ptr(), sp()))
    assert torch.allclose(out.cpu(), out_ref.detach(), rtol=1e-5, atol=1e-5)
    dy = torch.empty_like(yd); dw = torch.empty(64, device=d); dbias = torch.empty(64, device=d)
    wsp = ws()
    dod = d_out.to(d)
    _lib.check(lib.adn_head1x1_backward(yd.data_ptr(), dod.data_ptr(), wd.data_ptr(), pixels, dy.data_ptr(), dw.data_ptr(),
                                        dbias.data_ptr(), wsp.data_ptr(), sp()))
    assert rel(dw.cpu(), wt.grad.reshape(-1)) < 1e-4 and abs(float(dbias[0]) - float(b.grad)) < 1e-3 * abs(float(b.grad)) + 1e-4
    assert rel(nchw(dy), yr.grad) < 5e-3


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 16, 16, 64, 64), (2, 8, 8, 128, 128), (1, 16, 4, 64, 256), (3, 24, 8, 256, 64), (1, 4, 4, 512, 1024)])
def test_conv3x3_wgrad_matches_autograd(n, h, w, ci, co):
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    x = bf(torch.randn(n, ci, h, w, generator=g))
    dz = bf(torch.randn(n, co, h, w, generator=g))
    wt = torch.zeros(co, ci, 3, 3, requires_grad=True)
    F.conv2d(x, wt, padding=1).backward(dz)
    dw = torch.zeros((co, ci, 3, 3), dtype=torch.float32, device=dev())
    dzd, xd = nhwc(dz), nhwc(x)
    wk = wgws()
    _lib.check(lib.adn_conv3x3_wgrad_f32(dzd.data_ptr(), co, xd.data_ptr(), ci, h, w, n, h, w, dw.data_ptr(), 0, ci, wk.data_ptr(), sp()))
    assert rel(dw.cpu(), wt.grad) < 2e-3


def test_conv3x3_wgrad_concat_slice_and_padded_source():
    """decoder conv: input = cat([skip, F.pad(up)]) (model.py:44-49): two launches into column slices of one gradient."""
    lib = _lib.load()
    n, h, w, c = 2, 9, 17, 64
    g = torch.Generator().manual_seed(6)
    skip = bf(torch.randn(n, c, h, w, generator=g))
    up = bf(torch.randn(n, c, h - 1, w - 1, generator=g))
    dz = bf(torch.randn(n, 128, h, w, generator=g))
    wt = torch.zeros(128, 2 * c, 3, 3, requires_grad=True)
    F.conv2d(torch.cat([skip, F.pad(up, [0, 1, 0, 1])], dim=1), wt, padding=1).backward(dz)
    dw = torch.zeros((128, 2 * c, 3, 3), dtype=torch.float32, device=dev())
    dzd, skd, upd = nhwc(dz), nhwc(skip), nhwc(up)
    wk = wgws()
    _lib.check(lib.adn_conv3x3_wgrad_f32(dzd.data_ptr(), 128, skd.data_ptr(), c, h, w, n, h, w, dw.data_ptr(), 0, 2 * c, wk.data_ptr(), sp()))
    _lib.check(lib.adn_conv3x3_wgrad_f32(dzd.data_ptr(), 128, upd.data_ptr(), c, h - 1, w - 1, n, h, w, dw.data_ptr(), c, 2 * c, wk.data_ptr(), sp()))
    assert rel(dw.cpu(), wt.grad) < 2e-3


def test_conv3x3_c1_wgrad():
    lib = _lib.load()
    n, h, w = 3, 16, 24
    g = torch.Generator().manual_seed(7)
    x = torch.rand(n, 1, h, w, generator=g)
    dz = bf(torch.randn(n, 64, h, w, generator=g))
    wt = torch.zeros(64, 1, 3, 3, requires_grad=True)
    F.conv2d(x, wt, padding=1).backward(dz)
    dw = torch.empty((64, 1, 3, 3), dtype=torch.float32, device=dev())
    wsp = ws()
    dzd, xd = nhwc(dz), x.to(dev())
    _lib.check(lib.adn_conv3x3_c1_wgrad_f32(dzd.data_ptr(), xd.data_ptr(), n, h, w, dw.data_ptr(), wsp.data_ptr(), sp()))
    assert rel(dw.cpu(), wt.grad) < 1e-4


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (1, 8, 24, 128)])
def test_conv3x3_dgrad_via_forward_kernel(n, h, w, c):
    lib = _lib.load()
    co = 2 * c
    g = torch.Generator().manual_seed(8)
    wt = bf(torch.randn(co, c, 3, 3, generator=g) * 0.05)
    x = torch.zeros(n, c, h, w, requires_grad=True)
    dz = bf(torch.randn(n, co, h, w, generator=g))
    F.conv2d(x, wt, padding=1).backward(dz)
    wp = torch.empty((c, 9, co), dtype=torch.bfloat16, device=dev())
    wtd, dzd = wt.to(dev()), nhwc(dz)
    _lib.check(lib.adn_pack_conv3x3_dgrad_weight_bf16(wtd.data_ptr(), co, c, wp.data_ptr(), sp()))
    ones, zeros = torch.ones(c, device=dev()), torch.zeros(c, device=dev())
    dx = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=dev())
    _lib.check(lib.adn_conv3x3_affine_bf16(dzd.data_ptr(), co, 0, 0, 0, 0, n, h, w, wp.data_ptr(), c, ones.data_ptr(), zeros.data_ptr(), 0,
                                           dx.data_ptr(), sp()))
    assert rel(nchw(dx), x.grad) < 6e-3


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 8, 8, 128, 64), (1, 4, 12, 256, 128), (2, 2, 2, 1024, 512)])
def test_convt2x2_backward(n, h, w, ci, co):
    lib = _lib.load()
    g = torch.Generator().manual_seed(9)
    x = bf(torch.randn(n, ci, h, w, generator=g))
    wt = bf(torch.randn(ci, co, 2, 2, generator=g) * (1.0 / ci) ** 0.5)
    xr, wr = x.clone().requires_grad_(), wt.clone().requires_grad_()
    bias = torch.zeros(co, requires_grad=True)
    d_up = bf(torch.randn(n, co, 2 * h, 2 * w, generator=g))
    F.conv_transpose2d(xr, wr, bias, stride=2).backward(d_up)
    d = dev()
    # d_up lives in channels [co, 2co) of a wider tensor, as the decoder conv's data gradient does
    wide = torch.zeros((n, 2 * h, 2 * w, 2 * co), dtype=torch.bfloat16, device=d)
    wide[..., co:] = nhwc(d_up)
    dw = torch.zeros((ci, co, 2, 2), dtype=torch.float32, device=d)
    xd, wtd = nhwc(x), wt.to(d)
    wk = wgws()
    _lib.check(lib.adn_convt2x2_wgrad_f32(xd.data_ptr(), ci, wide.data_ptr(), 2 * co, co, co, n, h, w, dw.data_ptr(), wk.data_ptr(), sp()))
    assert rel(dw.cpu(), wr.grad) < 2e-3
    wp = torch.empty((ci, 4 * co), dtype=torch.bfloat16, device=d)
    _lib.check(lib.adn_pack_convt2x2_dgrad_weight_bf16(wtd.data_ptr(), ci, co, wp.data_ptr(), sp()))
    dx = torch.empty((n, h, w, ci), dtype=torch.bfloat16, device=d)
    _lib.check(lib.adn_convt2x2_dgrad_bf16(wide.data_ptr(), 2 * co, co, co, n, h, w, wp.data_ptr(), ci, dx.data_ptr(), sp()))
    assert rel(nchw(dx), xr.grad) < 6e-3
    dbias = torch.empty(co, dtype=torch.float32, device=d)
    wsp = ws()
    _lib.check(lib.adn_channel_sum_f32(wide.data_ptr() + 2 * co, 2 * co, n * 4 * h * w, co, dbias.data_ptr(), wsp.data_ptr(), sp()))
    assert rel(dbias.cpu(), bias.grad) < 1e-4


@pytest.mark.parametrize("b,f,t", [(2, 256, 64), (3, 32, 40), (1, 257, 188)])
def test_loss_backward_matches_autograd(b, f, t):
    lib = _lib.load()
    g = torch.Generator().manual_seed(10)
    pred = torch.rand(b, 1, f, t, generator=g).requires_grad_()
    target = torch.rand(b, 1, f, t, generator=g)
    total, *_ = loss_oracle.combined_loss(pred, target)
    total.backward()
    d = dev()
    fb = mel_filterbank().to(d)
    need = int(lib.adn_loss_backward_workspace_bytes(b, f, t))
    wsp = torch.empty(need, dtype=torch.uint8, device=d)
    dp = torch.empty((b, 1, f, t), dtype=torch.float32, device=d)
    pd, td = pred.detach().to(d), target.to(d)
    _lib.check(lib.adn_combined_loss_backward_f32(pd.data_ptr(), td.data_ptr(), b, f, t, fb.data_ptr(), 0.4, 0.4, 0.2,
                                                  wsp.data_ptr(), dp.data_ptr(), sp()))
    ref = pred.grad
    assert rel(dp.cpu(), ref) < 2e-3
    # the envelope part alone (subtract the L1 term): same tolerance on its own scale
    l1 = 0.2 * torch.sign(pred.detach() - target) / pred.numel()
    assert rel(dp.cpu() - l1, ref - l1) < 5e-3


def test_grad_norm_and_adamw_match_torch():
    lib = _lib.load()
    g = torch.Generator().manual_seed(11)
    n = 100003
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * s for s in (3.0, 0.01, 1.0)]
    pr = p0.clone().requires_grad_()
    opt = torch.optim.AdamW([pr], lr=1e-4)
    d = dev()
    p = p0.to(d); m = torch.zeros(n, device=d); v = torch.zeros(n, device=d)
    nc = torch.empty(2, device=d)
    wsp = ws()
    for step, gr in enumerate(grads, 1):
        pr.grad = gr.clone()
        norm_ref = torch.nn.utils.clip_grad_norm_([pr], 1.0)
        opt.step()
        gd = gr.to(d)
        _lib.check(lib.adn_grad_norm_f32(gd.data_ptr(), n, 1.0, nc.data_ptr(), wsp.data_ptr(), sp()))
        assert abs(float(nc[0]) - float(norm_ref)) <= 1e-5 * float(norm_ref)
        _lib.check(lib.adn_adamw_step_f32(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, nc.data_ptr(), 1e-4, 0.9, 0.999, 1e-8, 0.01,
                                          step, sp()))
        assert float((p.cpu() - pr.detach()).abs().max()) < 5e-7


def test_pack_table_equals_individual_packs():
    """adn_pack_weights_table_bf16 (one launch for every layer, tiled through shared memory) writes exactly what the four
    per-tensor pack entry points write."""
    import numpy as np
    lib = _lib.load()
    d = dev()
    g = torch.Generator().manual_seed(12)
    shapes = [(0, 128, 64), (0, 64, 192), (1, 128, 64), (0, 256, 256)]           # (kind, c_out, c_in)
    rec = np.dtype([("w", np.uint64), ("fwd", np.uint64), ("dgrad", np.uint64), ("c_out", np.int32), ("c_in", np.int32),
                    ("kind", np.int32), ("pad", np.int32)])
    keep, rows, refs = [], [], []
    for kind, co, ci in shapes:
        w = (torch.randn((co, ci, 3, 3) if kind == 0 else (ci, co, 2, 2), generator=g)).to(d)
        n = w.numel()
        fwd, dg = torch.zeros(n, dtype=torch.bfloat16, device=d), torch.zeros(n, dtype=torch.bfloat16, device=d)
        rf, rd = torch.zeros_like(fwd), torch.zeros_like(dg)
        if kind == 0:
            _lib.check(lib.adn_pack_conv3x3_weight_bf16(w.data_ptr(), co, ci, rf.data_ptr(), sp()))
            _lib.check(lib.adn_pack_conv3x3_dgrad_weight_bf16(w.data_ptr(), co, ci, rd.data_ptr(), sp()))
        else:
            _lib.check(lib.adn_pack_convt2x2_weight_bf16(w.data_ptr(), ci, co, rf.data_ptr(), sp()))
            _lib.check(lib.adn_pack_convt2x2_dgrad_weight_bf16(w.data_ptr(), ci, co, rd.data_ptr(), sp()))
        rows.append((w.data_ptr(), fwd.data_ptr(), dg.data_ptr(), co, ci, kind, 0))
        keep.append((w, fwd, dg)); refs.append((rf, rd))
    table = torch.from_numpy(np.array(rows, dtype=rec).view(np.uint8).copy()).to(d)
    _lib.check(lib.adn_pack_weights_table_bf16(table.data_ptr(), len(rows), sp()))
    for (w, fwd, dg), (rf, rd) in zip(keep, refs):
        assert torch.equal(fwd, rf) and torch.equal(dg, rd)
    # the one-dimensional grid (what the engine launches): `pad` = blocks per entry; and the forward / data-gradient selection
    blocks = [(co // 32) * (ci // 32) if kind == 0 else 3 for kind, co, ci in shapes]
    rows2 = [r[:6] + (b,) for r, b in zip(rows, blocks)]
    table2 = torch.from_numpy(np.array(rows2, dtype=rec).view(np.uint8).copy()).to(d)
    for which in (1, 2, 3):
        for (w, fwd, dg) in keep:
            fwd.zero_(); dg.zero_()
        _lib.check(lib.adn_pack_weights_table_flat_bf16(table2.data_ptr(), len(rows2), sum(blocks), which, sp()))
        for (w, fwd, dg), (rf, rd) in zip(keep, refs):
            assert torch.equal(fwd, rf if which & 1 else torch.zeros_like(rf))
            assert torch.equal(dg, rd if which & 2 else torch.zeros_like(rd))

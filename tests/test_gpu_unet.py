"""Parity of the tcgen05 UNet forward (through the C ABI) against the reference's own model.py outputs
(tests/golden/unet_*.npz, made by oracle/make_golden.py from /root/reference/code/model.py) and the fp32 oracle.

Tolerance (BASELINE.json north_star): bf16 model output within 1e-2 relative error -- norm-wise ||d|| / ||ref||
 -- and output SNR within 0.05 dB."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from audiodenoiser_b200 import _lib
from audiodenoiser_b200.checkpoint import seeded_state_dict
from audiodenoiser_b200.model import UNet
from oracle import unet_oracle

pytestmark = pytest.mark.gpu
UNET_TOL = 1e-2
SNR_TOL_DB = 0.05


def dev():
    return torch.device("cuda", 0)


def nrel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def to_nhwc_bf16(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def from_nhwc(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


@pytest.fixture(scope="module")
def net():
    m = UNet().eval()
    m.load_state_dict(seeded_state_dict(7))
    return m


@pytest.mark.parametrize("n,h,w", [(1, 1, 1), (2, 3, 7), (3, 17, 64), (1, 5, 523), (2, 3, 1034), (1, 2, 12001)])
@pytest.mark.parametrize("relu", [1, 0])
def test_conv3x3_first_layer(n, h, w, relu):
    """model.py:9-11 with one input channel (fp32 math, folded affine, optional ReLU -> NHWC bf16): short rows (4-pixel runs,
    staged), wide ragged rows (8-pixel runs, staged) and rows too wide to stage (direct path)."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(n * 1000 + h * 10 + w)
    x = torch.randn(n, 1, h, w, generator=g)
    wt = torch.randn(64, 1, 3, 3, generator=g) * 0.3
    sc = torch.rand(64, generator=g) + 0.5
    sh = torch.randn(64, generator=g) * 0.2
    ref = F.conv2d(x, wt, padding=1) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    if relu:
        ref = F.relu(ref)
    xd, wd, scd, shd = x.to(dev()).contiguous(), wt.to(dev()).contiguous(), sc.to(dev()), sh.to(dev())
    out = torch.full((n, h, w, 64), float("nan"), dtype=torch.bfloat16, device=dev())
    _lib.check(lib.adn_conv3x3_c1_affine_bf16(xd.data_ptr(), n, h, w, wd.data_ptr(), scd.data_ptr(), shd.data_ptr(), relu,
                                              out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = from_nhwc(out.cpu())
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()        # one bf16 rounding of an fp32 result
    assert nrel(got.numpy(), ref.numpy()) < 4e-3


@pytest.mark.parametrize("n,c0,c1,co,h,w", [(2, 64, 0, 64, 20, 19), (1, 64, 0, 128, 37, 26), (2, 128, 0, 256, 9, 6),
                                            (1, 256, 256, 256, 16, 11), (2, 64, 64, 64, 33, 28), (1, 512, 0, 1024, 4, 3),
                                            (1, 512, 512, 512, 5, 7), (3, 64, 0, 64, 257, 188), (1, 64, 0, 64, 1, 1),
                                            (2, 128, 0, 256, 64, 157), (2, 128, 128, 512, 50, 163)])   # enough tiles for the CTA-pair kernels (256-wide, 1 and 2 n-blocks)
def test_conv3x3_bn_relu(n, c0, c1, co, h, w):
    """One implicit-GEMM conv against F.conv2d on the same bf16-rounded operands: only accumulation order and the
    final bf16 store differ, so the bound is tight (output rounding 2^-9 max-wise)."""
    lib = _lib.load(); s = _lib.stream_ptr()
    g = torch.Generator().manual_seed(1000 * h + w)
    ci = c0 + c1
    x0 = bf16_round(torch.randn(n, c0, h, w, generator=g))
    h1, w1 = (max(h - (h % 2), 1), max(w - (w % 2), 1)) if c1 else (0, 0)
    x1 = bf16_round(torch.randn(n, c1, h1, w1, generator=g)) if c1 else None
    wt = bf16_round(torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (9 * ci)) ** 0.5)
    scale = 0.5 + torch.rand(co, generator=g); shift = 0.1 * torch.randn(co, generator=g)
    xin = x0 if x1 is None else torch.cat([x0, F.pad(x1, [0, w - w1, 0, h - h1])], 1)
    ref = F.relu(F.conv2d(xin, wt, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    wp = torch.empty((co, 9, ci), dtype=torch.bfloat16, device=dev())
    wd = wt.to(dev()).contiguous()
    _lib.check(lib.adn_pack_conv3x3_weight_bf16(wd.data_ptr(), co, ci, wp.data_ptr(), s))
    a0 = to_nhwc_bf16(x0).to(dev()); a1 = to_nhwc_bf16(x1).to(dev()) if c1 else None
    out = torch.zeros((n, h, w, co), dtype=torch.bfloat16, device=dev())
    do_pool = h >= 2 and w >= 2
    pool = torch.zeros((n, h // 2, w // 2, co), dtype=torch.bfloat16, device=dev()) if do_pool else None
    sc, sh = scale.to(dev()), shift.to(dev())
    _lib.check(lib.adn_conv3x3_bn_relu_bf16(a0.data_ptr(), c0, a1.data_ptr() if c1 else 0, c1, h1, w1, n, h, w, wp.data_ptr(), co,
                                            sc.data_ptr(), sh.data_ptr(), out.data_ptr(), pool.data_ptr() if do_pool else 0, s))
    torch.cuda.synchronize()
    got = from_nhwc(out.cpu())
    assert float((got - ref).abs().max()) <= 4e-3 * float(ref.abs().max())
    assert nrel(got, ref) <= 2.5e-3
    if do_pool:
        assert torch.equal(from_nhwc(pool.cpu()), F.max_pool2d(got, 2))         # pooling of the stored values is exact


@pytest.mark.parametrize("co", [128, 256, 512])
def test_conv3x3_pair_and_pitch_variants_bit_equal(co):
    """conv_halo's tuning hooks: CTA pairs off / 128-wide only / 128- and 256-wide (adn__conv_pair_mode 0 / 1 / 3), the
    padded 16-pixel halo pitch (adn__conv_halo_pitch) and the resident / streamed weight choice (adn__conv_halo_tune) are schedules of the same arithmetic: outputs must be bit-identical."""
    import ctypes
    lib = _lib.load(); s = _lib.stream_ptr()
    for f in (lib.adn__conv_pair_mode, lib.adn__conv_halo_pitch, lib.adn__conv_halo_tune):
        f.argtypes = [ctypes.c_int]; f.restype = None
    n, ci, h, w = 2, 128, 66, 157
    g = torch.Generator().manual_seed(co)
    x = to_nhwc_bf16(torch.randn(n, ci, h, w, generator=g)).to(dev())
    wt = (torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (9 * ci)) ** 0.5).to(dev())
    wp = torch.empty((co, 9, ci), dtype=torch.bfloat16, device=dev())
    _lib.check(lib.adn_pack_conv3x3_weight_bf16(wt.data_ptr(), co, ci, wp.data_ptr(), s))
    sc = (0.5 + torch.rand(co, generator=g)).to(dev()); sh = (0.1 * torch.randn(co, generator=g)).to(dev())
    outs = []
    try:
        for pair, pitch, tune in ((0, 10, 2), (1, 10, 2), (3, 10, 2), (3, 16, 0), (0, 16, 0), (3, 10, 0), (1, 10, 1), (3, 10, 6), (0, 10, 6)):
            lib.adn__conv_pair_mode(pair); lib.adn__conv_halo_pitch(pitch); lib.adn__conv_halo_tune(tune)
            out = torch.zeros((n, h, w, co), dtype=torch.bfloat16, device=dev())
            pool = torch.zeros((n, h // 2, w // 2, co), dtype=torch.bfloat16, device=dev())
            _lib.check(lib.adn_conv3x3_bn_relu_bf16(x.data_ptr(), ci, 0, 0, 0, 0, n, h, w, wp.data_ptr(), co, sc.data_ptr(), sh.data_ptr(),
                                                    out.data_ptr(), pool.data_ptr(), s))
            torch.cuda.synchronize()
            outs.append((out.cpu(), pool.cpu()))
    finally:
        lib.adn__conv_pair_mode(3); lib.adn__conv_halo_pitch(10); lib.adn__conv_halo_tune(2)
    assert outs[0][0].float().abs().max() > 0
    for o, p in outs[1:]:
        assert torch.equal(o, outs[0][0]) and torch.equal(p, outs[0][1])


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("n,c0,c1,h,w,relu", [(3, 64, 0, 257, 188, 1), (2, 64, 64, 61, 90, 0), (5, 128, 0, 64, 256, 1)])
def test_conv3x3_64_output_channels_all_kernels(mode, n, c0, c1, h, w, relu):
    """The 64-output-channel layers have three implementations behind one entry point (debug hook adn__conv_dx_mode): the
    halo-tile kernel (0), the kx-in-N kernel (1, default; 3 and 4 force four and two epilogue warp sets) and its CTA-pair variant (2).  All must agree with F.conv2d,
    with and without the activation, with the fused pool and across a concat."""
    import ctypes
    lib = _lib.load(); s = _lib.stream_ptr()
    lib.adn__conv_dx_mode.argtypes = [ctypes.c_int]; lib.adn__conv_dx_mode.restype = None
    g = torch.Generator().manual_seed(77 * h + w)
    ci = c0 + c1
    h1, w1 = (h - (h % 2), w - (w % 2)) if c1 else (0, 0)
    x0 = bf16_round(torch.randn(n, c0, h, w, generator=g))
    x1 = bf16_round(torch.randn(n, c1, h1, w1, generator=g)) if c1 else None
    wt = bf16_round(torch.randn(64, ci, 3, 3, generator=g) * (2.0 / (9 * ci)) ** 0.5)
    scale = 0.5 + torch.rand(64, generator=g); shift = 0.1 * torch.randn(64, generator=g)
    xin = x0 if x1 is None else torch.cat([x0, F.pad(x1, [0, w - w1, 0, h - h1])], 1)
    ref = F.conv2d(xin, wt, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if relu:
        ref = F.relu(ref)
    wp = torch.empty((64, 9, ci), dtype=torch.bfloat16, device=dev())
    wd = wt.to(dev()).contiguous()
    _lib.check(lib.adn_pack_conv3x3_weight_bf16(wd.data_ptr(), 64, ci, wp.data_ptr(), s))
    a0 = to_nhwc_bf16(x0).to(dev()); a1 = to_nhwc_bf16(x1).to(dev()) if c1 else None
    out = torch.zeros((n, h, w, 64), dtype=torch.bfloat16, device=dev())
    pool = torch.zeros((n, h // 2, w // 2, 64), dtype=torch.bfloat16, device=dev())
    sc, sh = scale.to(dev()), shift.to(dev())
    try:
        lib.adn__conv_dx_mode(mode)
        if relu:
            _lib.check(lib.adn_conv3x3_bn_relu_bf16(a0.data_ptr(), c0, a1.data_ptr() if c1 else 0, c1, h1, w1, n, h, w, wp.data_ptr(), 64,
                                                    sc.data_ptr(), sh.data_ptr(), out.data_ptr(), pool.data_ptr(), s))
        else:
            _lib.check(lib.adn_conv3x3_affine_bf16(a0.data_ptr(), c0, a1.data_ptr() if c1 else 0, c1, h1, w1, n, h, w, wp.data_ptr(), 64,
                                                   sc.data_ptr(), sh.data_ptr(), 0, out.data_ptr(), s))
        torch.cuda.synchronize()
    finally:
        lib.adn__conv_dx_mode(1)
    got = from_nhwc(out.cpu())
    assert float((got - ref).abs().max()) <= 4e-3 * float(ref.abs().max())
    assert nrel(got, ref) <= 2.5e-3
    if relu:
        assert torch.equal(from_nhwc(pool.cpu()), F.max_pool2d(got, 2))


@pytest.mark.parametrize("exact_weights", [True, False])
@pytest.mark.parametrize("n,c0,cl,co,hl,wl,h,w", [(2, 64, 128, 128, 8, 9, 17, 18), (1, 128, 256, 256, 5, 20, 10, 41), (3, 64, 128, 512, 17, 9, 35, 18),
                                                  (2, 128, 128, 128, 1, 1, 2, 2), (4, 64, 64, 256, 33, 70, 66, 141),
                                                  (1, 64, 128, 128, 20, 37, 41, 75), (4, 128, 256, 128, 16, 33, 32, 67)])
def test_conv3x3_upmerged_matches_convtranspose_pad_cat_conv(n, c0, cl, co, hl, wl, h, w, exact_weights):
    """UpSampleLayer.forward (model.py:41-50) up to the first ReLU -- ConvTranspose2d(k=2,s=2) + bias, F.pad to the skip's size,
    cat([skip, up]), Conv3x3 + folded BN + ReLU -- against the merged-weight kernel (csrc/conv_upm.cu), which never forms `up`.
    exact_weights: conv / ConvTranspose weights are small integers times a power of two, so the merged products are exact in
    bf16 and only accumulation order and the final bf16 store differ (tight bound, also proves the border-bias correction);
    otherwise the merged weights carry one bf16 rounding of their own (the two-layer path rounds `up` instead)."""
    lib = _lib.load(); s = _lib.stream_ptr()
    cup = cl // 2
    g = torch.Generator().manual_seed(h * 100 + w + co)
    skip = bf16_round(torch.randn(n, c0, h, w, generator=g))
    low = bf16_round(torch.randn(n, cl, hl, wl, generator=g))
    if exact_weights:
        w3 = torch.randint(-1, 2, (co, c0 + cup, 3, 3), generator=g).float() * 2.0 ** -5
        wt = torch.randint(-1, 2, (cl, cup, 2, 2), generator=g).float() * (torch.rand(cl, cup, 2, 2, generator=g) < 0.25).float() * 0.25
    else:
        w3 = torch.randn(co, c0 + cup, 3, 3, generator=g) * (2.0 / (9 * (c0 + cup))) ** 0.5
        wt = torch.randn(cl, cup, 2, 2, generator=g) * (1.0 / cl) ** 0.5
    bt = torch.randn(cup, generator=g) * 0.5
    scale = 0.5 + torch.rand(co, generator=g); shift = 0.1 * torch.randn(co, generator=g)
    up = F.conv_transpose2d(low.double(), wt.double(), bt.double(), stride=2)
    up = F.pad(up, [0, w - 2 * wl, 0, h - 2 * hl])
    w3r = w3 if exact_weights else torch.cat([bf16_round(w3[:, :c0]), w3[:, c0:]], 1)      # the skip half is rounded like any conv weight
    ref = F.relu(F.conv2d(torch.cat([skip.double(), up], 1), w3r.double(), padding=1) * scale.double().view(1, -1, 1, 1)
                 + shift.double().view(1, -1, 1, 1)).float()
    d = dev()
    wm = torch.empty((co, 9 * c0 + 16 * cl), dtype=torch.bfloat16, device=d)
    shift_m = torch.empty(co, device=d); wb = torch.empty((9, co), device=d)
    w3d, wtd, btd, sc, sh = (t.to(d).contiguous() for t in (w3, wt, bt, scale, shift))
    _lib.check(lib.adn_pack_upmerged_weight_bf16(w3d.data_ptr(), wtd.data_ptr(), btd.data_ptr(), sc.data_ptr(), sh.data_ptr(), co, c0, cup, cl,
                                                 wm.data_ptr(), shift_m.data_ptr(), wb.data_ptr(), s))
    a0 = to_nhwc_bf16(skip).to(d); a1 = to_nhwc_bf16(low).to(d)
    out = torch.full((n, h, w, co), float("nan"), dtype=torch.bfloat16, device=d)
    _lib.check(lib.adn_conv3x3_upmerged_bn_relu_bf16(a0.data_ptr(), c0, a1.data_ptr(), cl, hl, wl, n, h, w, wm.data_ptr(), co,
                                                     sc.data_ptr(), shift_m.data_ptr(), wb.data_ptr(), out.data_ptr(), s))
    torch.cuda.synchronize()
    outs = [out]
    if co == 128:                 # the two-classes-per-tile kernel (conv3x3_upm2_kernel) on the same merged weights
        import ctypes
        lib.adn_upmerged_pair_weight_elems.restype = ctypes.c_int64
        bsh = torch.empty(int(lib.adn_upmerged_pair_weight_elems(co, c0, cl, 0)), dtype=torch.bfloat16, device=d)
        b1 = torch.empty(int(lib.adn_upmerged_pair_weight_elems(co, c0, cl, 1)), dtype=torch.bfloat16, device=d)
        _lib.check(lib.adn_pack_upmerged_pair_weight_bf16(wm.data_ptr(), co, c0, cl, bsh.data_ptr(), b1.data_ptr(), s))
        out2 = torch.full((n, h, w, co), float("nan"), dtype=torch.bfloat16, device=d)
        _lib.check(lib.adn_conv3x3_upmerged_pair_bn_relu_bf16(a0.data_ptr(), c0, a1.data_ptr(), cl, hl, wl, n, h, w, bsh.data_ptr(), b1.data_ptr(),
                                                              co, sc.data_ptr(), shift_m.data_ptr(), wb.data_ptr(), out2.data_ptr(), s))
        torch.cuda.synchronize()
        outs.append(out2)
    for o in outs:
        got = from_nhwc(o.cpu())
        assert torch.isfinite(got).all()
        if exact_weights:
            assert float((got - ref).abs().max()) <= 4e-3 * float(ref.abs().max())
            assert nrel(got, ref) <= 2.5e-3
        else:
            assert float((got - ref).abs().max()) <= 1.2e-2 * float(ref.abs().max())
            assert nrel(got, ref) <= 4e-3


@pytest.mark.parametrize("n,c0,h,w", [(2, 128, 17, 18), (1, 64, 40, 75), (4, 128, 32, 67), (3, 256, 2, 2), (2, 128, 128, 517)])
def test_conv3x3_parity_class_kernel_without_low_tensor(n, c0, h, w):
    """conv3x3_upm2_kernel with cl = 0 is a plain Conv3x3 + BN + ReLU with 128 output channels computed per row-parity class (two
    column-parity classes per tile): against F.conv2d on the same bf16 operands, odd and even sizes, one and several tiles."""
    import ctypes
    lib = _lib.load(); s = _lib.stream_ptr()
    lib.adn_upmerged_pair_weight_elems.restype = ctypes.c_int64
    co = 128
    g = torch.Generator().manual_seed(h * 7 + w)
    x = bf16_round(torch.randn(n, c0, h, w, generator=g))
    wt = bf16_round(torch.randn(co, c0, 3, 3, generator=g) * (2.0 / (9 * c0)) ** 0.5)
    scale = 0.5 + torch.rand(co, generator=g); shift = 0.1 * torch.randn(co, generator=g)
    ref = F.relu(F.conv2d(x, wt, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    d = dev()
    wp = torch.empty((co, 9, c0), dtype=torch.bfloat16, device=d)
    wd = wt.to(d).contiguous()
    _lib.check(lib.adn_pack_conv3x3_weight_bf16(wd.data_ptr(), co, c0, wp.data_ptr(), s))
    bsh = torch.empty(int(lib.adn_upmerged_pair_weight_elems(co, c0, 0, 0)), dtype=torch.bfloat16, device=d)
    b1 = torch.empty(int(lib.adn_upmerged_pair_weight_elems(co, c0, 0, 1)), dtype=torch.bfloat16, device=d)
    _lib.check(lib.adn_pack_upmerged_pair_weight_bf16(wp.data_ptr(), co, c0, 0, bsh.data_ptr(), b1.data_ptr(), s))
    a0 = to_nhwc_bf16(x).to(d)
    out = torch.full((n, h, w, co), float("nan"), dtype=torch.bfloat16, device=d)
    sc, sh = scale.to(d), shift.to(d)
    _lib.check(lib.adn_conv3x3_upmerged_pair_bn_relu_bf16(a0.data_ptr(), c0, 0, 0, 0, 0, n, h, w, bsh.data_ptr(), b1.data_ptr(), co,
                                                          sc.data_ptr(), sh.data_ptr(), 0, out.data_ptr(), s))
    torch.cuda.synchronize()
    got = from_nhwc(out.cpu())
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= 4e-3 * float(ref.abs().max())
    assert nrel(got, ref) <= 2.5e-3
    # the fused-pool variant: same conv output bit for bit, pooled output = exact 2x2 floor pooling of the stored values
    out2 = torch.full((n, h, w, co), float("nan"), dtype=torch.bfloat16, device=d)
    pool = torch.full((n, h // 2, w // 2, co), float("nan"), dtype=torch.bfloat16, device=d)
    _lib.check(lib.adn_conv3x3_pair_bn_relu_pool_bf16(a0.data_ptr(), c0, n, h, w, bsh.data_ptr(), b1.data_ptr(), co, sc.data_ptr(), sh.data_ptr(),
                                                      out2.data_ptr(), pool.data_ptr(), s))
    torch.cuda.synchronize()
    assert torch.equal(out2.cpu(), out.cpu())
    assert torch.equal(from_nhwc(pool.cpu()), F.max_pool2d(got, 2))


@pytest.mark.parametrize("n,ci,co,h,w", [(2, 128, 64, 9, 7), (1, 1024, 512, 2, 3), (2, 256, 128, 16, 11), (1, 512, 256, 1, 1)])
def test_convt2x2(n, ci, co, h, w):
    lib = _lib.load(); s = _lib.stream_ptr()
    g = torch.Generator().manual_seed(ci + h)
    x = bf16_round(torch.randn(n, ci, h, w, generator=g))
    wt = bf16_round(torch.randn(ci, co, 2, 2, generator=g) * (1.0 / ci) ** 0.5)
    b = 0.1 * torch.randn(co, generator=g)
    ref = F.conv_transpose2d(x, wt, b, stride=2)
    wp = torch.empty((4, co, ci), dtype=torch.bfloat16, device=dev())
    wd = wt.to(dev()).contiguous()
    _lib.check(lib.adn_pack_convt2x2_weight_bf16(wd.data_ptr(), ci, co, wp.data_ptr(), s))
    a = to_nhwc_bf16(x).to(dev()); bd = b.to(dev())
    out = torch.zeros((n, 2 * h, 2 * w, co), dtype=torch.bfloat16, device=dev())
    _lib.check(lib.adn_convt2x2_bf16(a.data_ptr(), ci, n, h, w, wp.data_ptr(), co, bd.data_ptr(), out.data_ptr(), s))
    torch.cuda.synchronize()
    got = from_nhwc(out.cpu())
    assert float((got - ref).abs().max()) <= 4e-3 * float(ref.abs().max())


@pytest.mark.parametrize("name", ["small", "train", "test"])
def test_unet_matches_reference_fixture(net, golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"unet_{name}.npz"))
    y = net(torch.from_numpy(z["x"]).to(dev()))
    assert y.shape == z["y"].shape and y.dtype == torch.float32 and y.is_cuda
    y = y.cpu().numpy()
    assert nrel(y, z["y"]) <= UNET_TOL
    # output SNR of the "denoised" magnitude against a target, reference-fp32 model vs this build
    clean = 0.5 * z["x"]
    snr = lambda d: 10.0 * np.log10(np.sum(clean.astype(np.float64) ** 2) / np.sum((clean.astype(np.float64) - d) ** 2))
    assert abs(snr(y) - snr(z["y"])) <= SNR_TOL_DB


def _ckpt(name):
    if name == "default":                      # the reference's own default initialisation: the drop-in constructor reproduces it
        torch.manual_seed(5)
        return {k: v.detach().clone() for k, v in UNet().state_dict().items()}
    return seeded_state_dict(int(name))


@pytest.mark.parametrize("fixture,ckpt", [("unet_B", "7"), ("unet_ckpt_3", "3"), ("unet_ckpt_11", "11"), ("unet_ckpt_default", "default")])
def test_unet_matches_reference_on_more_checkpoints_and_the_benchmark_shape(golden_dir, fixture, ckpt):
    """the full UNet at the BENCHMARK shape (1,1,257,1034) and on three more checkpoints (two seeded ones
    and the reference's default init with identity BatchNorm statistics), each against outputs of the reference's own model.py
    (oracle/make_golden_r2.py).  Bounds: the north_star's 1e-2 norm-wise, and the kernels' error must sit at the level of the bf16
    recipe's own error as the fp32 emulation of the same roundings measures it (oracle unet_forward(emulate_bf16=True)).  An
    element-wise match with the emulation is not expected: rounding decisions decorrelate after a few layers (Numerics)."""
    z = np.load(os.path.join(golden_dir, f"{fixture}.npz"))
    x = torch.from_numpy(z["x"].astype(np.float32))
    sd = _ckpt(ckpt)
    m = UNet().eval(); m.load_state_dict(sd)
    y = m(x.to(dev())).cpu().numpy()
    ref = z["y"]
    assert y.shape == ref.shape
    e_ref = nrel(y, ref)
    e_recipe = nrel(unet_oracle.unet_forward(sd, x, emulate_bf16=True).numpy(), ref)
    print(f"{fixture}: vs reference {e_ref:.4%}, bf16-recipe emulation vs reference {e_recipe:.4%}")
    assert e_ref <= UNET_TOL
    assert e_ref <= 1.1 * e_recipe + 5e-4
    clean = 0.5 * x.numpy()
    snr = lambda d: 10.0 * np.log10(np.sum(clean.astype(np.float64) ** 2) / np.sum((clean.astype(np.float64) - d) ** 2))
    assert abs(snr(y) - snr(ref)) <= SNR_TOL_DB


def test_unet_intermediate_levels(net):
    """Per-level parity against the oracle's intermediates: localises an error to a layer."""
    sd = seeded_state_dict(7)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet_small.npz"))
    x = torch.from_numpy(z["x"])
    _, inter = unet_oracle.unet_forward(sd, x, return_intermediates=True)
    net(x.to(dev())); torch.cuda.synchronize()
    ws = list(net._ws.values())[0]
    names = {"down1": "s0", "down2": "s1", "down3": "s2", "down4": "s3", "bottle": "s4", "up1": "ub3", "up2": "ub2", "up3": "ub1"}
    for k, b in names.items():
        assert nrel(from_nhwc(ws[b].cpu()).numpy(), inter[k].numpy()) <= UNET_TOL, k


def test_concat_order_and_pad_side(net):
    """model.py:49 cat([skip, up]) and model.py:44-47 pad right/bottom: zeroing the skip half of upconv4's first conv
    must match the oracle doing the same (odd 257 x 47 input exercises both pads)."""
    sd = seeded_state_dict(7)
    sd["upconv4.conv.double_conv.0.weight"][:, :64] = 0
    sd["upconv1.conv.double_conv.0.weight"][:, 512:] *= 2
    m = UNet().eval(); m.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(1, 1, 257, 47, generator=g) * 3
    ref = unet_oracle.unet_forward(sd, x).numpy()
    assert nrel(m(x.to(dev())).cpu().numpy(), ref) <= UNET_TOL


def test_state_dict_round_trip_and_repack(net):
    sd = seeded_state_dict(7)
    got = net.state_dict()
    assert list(got.keys()) == list(sd.keys()) and len(got) == 136
    for k in sd:
        assert got[k].dtype == sd[k].dtype and torch.equal(got[k].cpu(), sd[k])
    # a new checkpoint loaded into the same module must repack (no stale bf16 weights)
    m = UNet().eval(); m.load_state_dict(sd)
    x = torch.rand(1, 1, 64, 32) * 2
    y7 = m(x.to(dev())).cpu()
    sd3 = seeded_state_dict(3)
    m.load_state_dict(sd3)
    y3 = m(x.to(dev())).cpu().numpy()
    assert nrel(y3, unet_oracle.unet_forward(sd3, x).numpy()) <= UNET_TOL
    assert nrel(y7.numpy(), y3) > 0.05


def test_batch_larger_than_chunk_and_batch_invariance(net):
    """test.py:113 feeds the whole test set as one batch; results must not depend on the batch a clip sits in."""
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(70, 1, 32, 48, generator=g) * 2).to(dev())
    y = net(x)
    y1 = net(x[65:66])
    assert torch.equal(y[65:66], y1)


def test_host_tensor_in_host_tensor_out(net):
    """test.py:100,113: `model` stays on the CPU and is fed a CPU tensor; the result must be a CPU tensor ("on the input's
    device") with the same values the CUDA-tensor call yields -- host in / host out through the same kernels."""
    g = torch.Generator().manual_seed(21)
    x = torch.rand(3, 1, 257, 47, generator=g) * 2
    y_host = net(x)
    assert not y_host.is_cuda and y_host.dtype == torch.float32 and y_host.shape == x.shape
    assert torch.equal(y_host, net(x.to(dev())).cpu())
    big = torch.rand(2, 1, 257, 188, generator=g)                       # > 64 Ki elements: the pinned staging path
    assert torch.equal(net(big), net(big.to(dev())).cpu())
    # closeness to the fp32 reference on this input (uniform noise, not spectrogram-like): the bf16 recipe itself costs 1.05 % here
    # (oracle emulation), just outside the 1e-2 the north_star states for spectrogram inputs -- the kernels must not add to it
    sd = seeded_state_dict(7)
    ref = unet_oracle.unet_forward(sd, x).numpy()
    e_recipe = nrel(unet_oracle.unet_forward(sd, x, emulate_bf16=True).numpy(), ref)
    assert nrel(y_host.numpy(), ref) <= 1.1 * e_recipe + 5e-4 <= 1.3e-2


def test_input_validation(net):
    with pytest.raises(ValueError):
        net(torch.zeros(1, 64, 64))
    with pytest.raises(ValueError):
        net(torch.zeros(1, 2, 64, 64, device=dev()))
    with pytest.raises(ValueError):
        net(torch.zeros(1, 1, 8, 8, device=dev()))

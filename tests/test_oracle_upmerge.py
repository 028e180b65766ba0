"""CPU check of the algebra behind the merged ConvTranspose + conv kernel (csrc/conv_upm.cu): the parity-class formulation of
oracle/upmerge_oracle.py equals ConvTranspose2d -> F.pad -> cat -> Conv2d (model.py:41-49) to float64 round-off, for even and odd
sizes (F.pad with one missing row / column), including the border-pixel bias terms."""
import pytest
import torch

from oracle import upmerge_oracle as uo


@pytest.mark.parametrize("n,c0,cu,cl,hl,wl,dh,dw", [(1, 3, 2, 4, 3, 4, 0, 0), (2, 2, 3, 5, 2, 3, 1, 0), (1, 4, 2, 4, 3, 2, 0, 1), (1, 1, 1, 2, 1, 1, 1, 1),
                                                    (2, 3, 3, 6, 4, 5, 1, 1)])
def test_parity_class_formulation_equals_convtranspose_pad_cat_conv(n, c0, cu, cl, hl, wl, dh, dw):
    g = torch.Generator().manual_seed(100 * hl + wl + dh + 2 * dw)
    h, w = 2 * hl + dh, 2 * wl + dw
    skip = torch.randn(n, c0, h, w, generator=g)
    low = torch.randn(n, cl, hl, wl, generator=g)
    w3 = torch.randn(7, c0 + cu, 3, 3, generator=g)
    b3 = torch.randn(7, generator=g)
    wt = torch.randn(cl, cu, 2, 2, generator=g)
    bt = torch.randn(cu, generator=g)
    ref = uo.up_conv_reference(skip, low, w3, b3, wt, bt)
    got = uo.up_conv_merged(skip, low, w3, b3, wt, bt)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 1e-11 * max(1.0, float(ref.abs().max()))


def test_low_index_table():
    assert [uo.low_index(0, k) for k in range(3)] == [0, 1, 1]
    assert [uo.low_index(1, k) for k in range(3)] == [0, 0, 1]

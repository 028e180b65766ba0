#!/usr/bin/env python
"""Debug sweep of the MN-major UMMA descriptor strides of the wgrad kernel (LBO / SBO): prints the error of each combination."""
import ctypes, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodenoiser_b200 import _lib
lib = _lib.load()
setd = lib.adn__wgrad_set_desc
setd.argtypes = [ctypes.c_int, ctypes.c_int]; setd.restype = None
dev = torch.device("cuda", 0)
nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)
for (n, h, w, ci, co) in [(1, 8, 8, 64, 64), (2, 16, 16, 128, 128)]:
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, ci, h, w, generator=g).to(torch.bfloat16).float()
    dz = torch.randn(n, co, h, w, generator=g).to(torch.bfloat16).float()
    wt = torch.zeros(co, ci, 3, 3, requires_grad=True)
    F.conv2d(x, wt, padding=1).backward(dz)
    for lbo, sbo in [(8192, 1024), (1024, 8192), (8192, 128), (128, 1024), (1024, 1024), (2048, 1024)]:
        setd(lbo, sbo)
        dw = torch.zeros((co, ci, 3, 3), dtype=torch.float32, device=dev)
        dzd, xd = nhwc(dz), nhwc(x)
        st = lib.adn_conv3x3_wgrad_f32(dzd.data_ptr(), co, xd.data_ptr(), ci, h, w, n, h, w, dw.data_ptr(), 0, ci, _lib.stream_ptr())
        torch.cuda.synchronize()
        err = float((dw.cpu() - wt.grad).abs().max() / wt.grad.abs().max())
        print(f"shape {(n,h,w,ci,co)} lbo {lbo} sbo {sbo} status {st} rel err {err:.4g}", flush=True)
setd(8192, 1024)

#!/usr/bin/env python
"""Pure-write, pure-read and copy HBM bandwidth on this GPU (torch fill_ / cudaMemset / sum / copy_, 4 GiB buffers, CUDA events,
best of 10).  MEASURED_PEAKS.json's hbm_gbs is the COPY figure (read + write bytes); write-only kernels (first-layer conv, the
ConvTranspose layers, the pooled full-resolution conv) have to be judged against the write-only figure."""
import json

import torch

n = 1 << 30          # fp32 elements: 4 GiB
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def best(fn, reps=10):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for _ in range(2):
    a.fill_(1.0); b.copy_(a); a.sum()
torch.cuda.synchronize()
gb = n * 4 / 1e9
res = {
    "write_fill_GBps": gb / best(lambda: a.fill_(2.0)) * 1e3,
    "write_memset_GBps": gb / best(lambda: a.zero_()) * 1e3,
    "read_sum_GBps": gb / best(lambda: a.sum()) * 1e3,
    "copy_GBps_read_plus_write": 2 * gb / best(lambda: b.copy_(a)) * 1e3,
}
print(json.dumps(res))

"""Host logic of the clip-sharded path with world_size 2 and 3 over the gloo backend on CPU:
partitioning, ordered gather of unequal shards, exact statistics reduction."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiodenoiser_b200 import sharding


def test_shard_range_partitions_everything():
    for n in (0, 1, 5, 64, 65, 65536):
        for g in (1, 2, 3, 4, 8):
            rs = [sharding.shard_range(n, g, r) for r in range(g)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(hi - lo for lo, hi in rs) == -(-n // g)
            assert sum(sharding.shard_sizes(n, g)) == n
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_stats_from_sums():
    s = sharding.stats_from_sums([2.0, 100.0, 1.0, 8.0])
    assert s["l1"] == 0.25 and abs(s["snr_db"] - 20.0) < 1e-12 and s["count"] == 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_total * 3, dtype=torch.float32).reshape(n_total, 3)
        lo, hi = sharding.shard_range(n_total, world, rank)
        got = sharding.all_gather_rows(full[lo:hi].clone(), n_total)
        ok_gather = torch.equal(got, full)
        # asynchronous variant: the shard is copied into its slot first, so the source may be overwritten immediately
        per = -(-n_total // world)
        buf = torch.full((world * per, 3), -1.0)
        mine = full[lo:hi].clone()
        got2, work = sharding.all_gather_rows(mine, n_total, out=buf, async_op=True)
        mine.fill_(123.0)
        work.wait()
        ok_gather = ok_gather and torch.equal(got2, full)
        # statistics: per-rank partial sums -> exact global means
        pred = full[lo:hi] * 0.5
        d = full[lo:hi] - pred
        sums = torch.tensor([d.abs().sum().item(), (full[lo:hi] ** 2).sum().item(), (d ** 2).sum().item(), float(d.numel())],
                            dtype=torch.float64)
        sharding.all_reduce_sums(sums)
        st = sharding.stats_from_sums(sums)
        dd = full - 0.5 * full
        ok_stats = (abs(st["l1"] - dd.abs().double().mean().item()) < 1e-9 and st["count"] == full.numel()
                    and abs(st["snr_db"] - 10 * math.log10((full.double() ** 2).sum().item() / (dd.double() ** 2).sum().item())) < 1e-9)
        q.put((rank, ok_gather, ok_stats))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 5), (2, 8), (3, 7)])
def test_gather_and_reduce_over_gloo(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, True, True) for r in range(world)]

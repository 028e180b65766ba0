"""Data-parallel training host logic (BASELINE config 5) with world_size 2 over gloo on CPU: each rank takes the oracle's
gradient of ITS shard (per-replica BatchNorm, torch DDP semantics), sharding.average_gradients_ averages one flat buffer, and the
result equals the mean of the per-shard gradients computed in a single process."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiodenoiser_b200 import sharding
from audiodenoiser_b200.checkpoint import seeded_state_dict
from oracle.make_golden_train import batch
from oracle.train_oracle import TrainOracle


def _flat_grads(noisy, clean):
    torch.set_num_threads(2)
    orc = TrainOracle(seeded_state_dict(7))
    orc.forward_backward(noisy, clean)
    return torch.cat([p.grad.reshape(-1) for p in orc.params.values()])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        noisy, clean = batch(201, (4, 1, 32, 32))
        lo, hi = sharding.shard_range(4, world, rank)
        g = _flat_grads(noisy[lo:hi], clean[lo:hi])
        g2 = g.clone()
        sharding.average_gradients_(g)
        # the bucketed, asynchronous form TrainEngine.backward uses (three slices launched as they become final, waited later)
        cuts = [0, g2.numel() // 7, g2.numel() // 2, g2.numel()]
        handles = [sharding.average_gradients_async(g2[cuts[i]:cuts[i + 1]]) for i in (2, 1, 0)]
        for h in handles:
            h.wait()
        assert torch.equal(g2, g), "bucketed average differs from the single all-reduce"
        q.put((rank, g[::997].tolist()))           # plain floats: a tensor in a Queue needs the sender alive at receive time
    finally:
        dist.destroy_process_group()


def test_ddp_gradient_average_matches_per_replica_mean():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    noisy, clean = batch(201, (4, 1, 32, 32))
    ref = 0.5 * (_flat_grads(noisy[:2], clean[:2]) + _flat_grads(noisy[2:], clean[2:]))[::997]
    assert res[0] == res[1]                                    # every rank holds the same averaged gradient
    assert torch.allclose(torch.tensor(res[0]), ref, rtol=1e-5, atol=1e-7)


def test_average_gradients_is_identity_without_a_process_group():
    g = torch.arange(8, dtype=torch.float32)
    assert sharding.average_gradients_(g.clone()).equal(g)

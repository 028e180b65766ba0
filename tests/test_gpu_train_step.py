"""The whole training step (train.py:65-72) on the B200 kernels against the CPU oracle (oracle/train_oracle.py, itself pinned to
the reference's model.py / loss.py / AdamW by tests/test_oracle_train.py).

Tolerances.  BASELINE.json states 1e-2 for the bf16 *inference* output and nothing for training.  Two facts, both measured here:
  * in train() mode every BatchNorm re-normalises with batch statistics, which removes the channel mean and amplifies a relative
    perturbation by ~sqrt(1 + (mean/std)^2) per layer (~1.6x per encoder level with the seeded checkpoint): bf16 activations give
    6-10 % at the output of the 18 conv+BN layers for ANY bf16 implementation;
  * the reference loss is built from |.| terms, whose gradient is a sign(): an output perturbation of a few per cent flips many
    signs, so the gradient of the *actual* loss differs by tens of per cent between fp32 and any bf16 forward.
The yardstick is therefore measured in the same test: the same computation by stock torch on the GPU under autocast(bfloat16)
(cuDNN convolutions, bf16 activations) against the same fp32 CPU oracle.  Ours must be no worse than 1.5x torch-autocast's error.
The backward kernels themselves are additionally checked with a LINEAR loss (d_out fixed), where no sign can flip."""
import pytest
import torch

from audiodenoiser_b200.checkpoint import seeded_state_dict
from audiodenoiser_b200.model import UNet
from audiodenoiser_b200.training import TrainEngine
from oracle import loss_oracle
from oracle.make_golden_train import batch
from oracle.train_oracle import TrainOracle, split_state_dict, unet_forward_train

pytestmark = pytest.mark.gpu
DEAD = (".double_conv.0.bias", ".double_conv.3.bias")          # conv biases under train-mode BN: zero gradient (float noise in autograd)
DEV = "cuda:0"


def _engine(seed=7):
    net = UNet()
    net.load_state_dict(seeded_state_dict(seed))
    return TrainEngine(net, lr=1e-4, device=torch.device(DEV))


def _torch_step(noisy, clean, d_out=None, autocast=False):
    """The reference computation with stock torch: fp32 on the CPU (the oracle) or bf16 autocast on the GPU (the yardstick).
    d_out=None: the real loss; else the linear loss sum(out * d_out).  -> (out, {param: grad}) as fp32 CPU tensors."""
    p, b = split_state_dict(seeded_state_dict(7))
    if autocast:
        p = {k: v.detach().to(DEV).requires_grad_() for k, v in p.items()}
        b = {k: v.to(DEV) for k, v in b.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = unet_forward_train(p, b, noisy.to(DEV))
        out = out.float().cpu()                    # the (tiny) loss runs on the CPU oracle; its gradient flows back through .cpu()
    else:
        out = unet_forward_train(p, b, noisy)
    loss = loss_oracle.combined_loss(out, clean)[0] if d_out is None else (out * d_out).sum()
    loss.backward()
    return out.detach(), {k: v.grad.detach().float().cpu() for k, v in p.items()}


def _compare(eng, grads_ref, grads_ac, floor):
    worst = worst_ac = 0.0
    dots = [0.0, 0.0]; sq = [0.0, 0.0]; sq_ref = 0.0
    for k, g_ref in grads_ref.items():
        g = eng.gview(k).cpu()
        if k.endswith(DEAD):
            assert float(g.abs().max()) == 0.0
            continue
        e = float((g - g_ref).norm() / (g_ref.norm() + 1e-30))
        e_ac = float((grads_ac[k] - g_ref).norm() / (g_ref.norm() + 1e-30))
        worst, worst_ac = max(worst, e), max(worst_ac, e_ac)
        assert e <= max(floor, 1.5 * e_ac), f"{k}: gradient norm-wise error {e} (torch autocast bf16: {e_ac})"
        for j, t in enumerate((g, grads_ac[k])):
            dots[j] += float((t.double() * g_ref.double()).sum()); sq[j] += float(t.double().pow(2).sum())
        sq_ref += float(g_ref.double().pow(2).sum())
    cos, cos_ac = (dots[j] / (sq[j] ** 0.5 * sq_ref ** 0.5) for j in (0, 1))
    assert 1.0 - cos <= max(floor * floor, 1.5 * (1.0 - cos_ac)), f"gradient cosine similarity {cos} (torch autocast bf16: {cos_ac})"
    return worst, worst_ac, cos, cos_ac


@pytest.mark.parametrize("shape", [(2, 1, 32, 32), (4, 1, 256, 64)])
def test_forward_and_backward_match_oracle(shape):
    """train-mode forward + the whole backward with a fixed output gradient (no sign flips: isolates the kernels)."""
    noisy, clean = batch(101, shape)
    g = torch.Generator().manual_seed(5)
    d_out = torch.randn(shape, generator=g) / (shape[0] * shape[2] * shape[3])
    out_ref, grads_ref = _torch_step(noisy, clean, d_out)
    out_ac, grads_ac = _torch_step(noisy, clean, d_out, autocast=True)
    err_ac = float((out_ac - out_ref).norm() / out_ref.norm())
    eng = _engine()
    eng.zero_grad()
    out = eng.forward(noisy.to(DEV))
    err = float((out.cpu() - out_ref).norm() / out_ref.norm())
    assert err <= max(1e-2, 1.5 * err_ac), f"train-mode forward norm-wise error {err} (torch autocast bf16: {err_ac})"
    eng.backward(d_out.to(DEV))
    torch.cuda.synchronize()
    worst, worst_ac, cos, cos_ac = _compare(eng, grads_ref, grads_ac, floor=3e-2)
    print(f"linear loss {shape}: fwd err {err:.4f} (autocast {err_ac:.4f}); worst per-tensor grad err {worst:.4f} (autocast {worst_ac:.4f}); "
          f"cosine {cos:.6f} (autocast {cos_ac:.6f})")


@pytest.mark.parametrize("shape", [(2, 1, 32, 32), (4, 1, 256, 64)])
def test_train_step_matches_oracle(shape):
    """The literal step: real loss, clip, AdamW, BatchNorm running statistics."""
    torch.manual_seed(0)
    eng = _engine()
    orc = TrainOracle(seeded_state_dict(7), lr=1e-4, max_norm=1.0)
    noisy, clean = batch(101, shape)
    out_ref, losses_ref, grads_ref, norm_ref = orc.train_step(noisy, clean)
    _out_ac, grads_ac = _torch_step(noisy, clean, None, autocast=True)

    eng.zero_grad()
    out = eng.forward(noisy.to(DEV))
    losses, d_pred = eng.loss_and_grad(out, clean.to(DEV))
    assert torch.allclose(losses.cpu(), losses_ref, rtol=5e-2), (losses.cpu(), losses_ref)
    eng.backward(d_pred)
    torch.cuda.synchronize()
    worst, worst_ac, cos, cos_ac = _compare(eng, grads_ref, grads_ac, floor=3e-2)
    gn = float(eng.optimizer_step())
    gn_ac = sum(float(v.double().pow(2).sum()) for v in grads_ac.values()) ** 0.5
    # total gradient norm: a single scalar dominated by the deepest layers, whose train-mode BatchNorm normalises over as few as
    # 8 values at the small shape -- any bf16 implementation lands a few per cent off the fp32 oracle (torch autocast: 2.4 %,
    # this path 2.1 % .. 4.1 % across kernel revisions that only reorder fp32 sums).  Per-tensor errors are checked above.
    assert abs(gn - float(norm_ref)) <= max(5e-2 * float(norm_ref), 2.0 * abs(gn_ac - float(norm_ref)))

    sd_ref = orc.state_dict()
    for k, v in eng.model.state_dict().items():
        if k.endswith(("running_mean", "running_var")):
            ref = sd_ref[k]
            assert float((v.cpu() - ref).norm() / ref.norm()) < 2e-2, k
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(sd_ref[k])
    # one AdamW step moves every weight by ~lr * sign(g): the parameters stay within 2 lr of the oracle's, and they did move
    p0 = seeded_state_dict(7)
    for k, (o, n, shp) in eng.offsets.items():
        if k.endswith(DEAD):
            continue
        new, ref = eng.pview(k).cpu(), sd_ref[k]
        assert float((new - ref).abs().max()) <= 2.05e-4, k
        assert float((new - p0[k]).abs().max()) > 0.5e-4, k
    print(f"real loss {shape}: losses {losses.cpu().tolist()} vs {losses_ref.tolist()}; worst per-tensor grad err {worst:.4f} (autocast {worst_ac:.4f}); "
          f"cosine {cos:.6f} (autocast {cos_ac:.6f}); grad norm {gn:.4f} vs {float(norm_ref):.4f} (autocast {gn_ac:.4f})")


def test_two_steps_and_state_dict_round_trip():
    """Parameters stay visible through the module's state_dict (they alias the engine's flat buffer), the loss goes down on a
    repeated batch, and an eval() forward after training uses the updated weights / running statistics."""
    eng = _engine()
    noisy, clean = batch(102, (4, 1, 64, 32))
    noisy, clean = noisy.to(DEV), clean.to(DEV)
    first = eng.train_step(noisy, clean).cpu()
    for _ in range(7):
        last = eng.train_step(noisy, clean).cpu()
    assert float(last[0]) < float(first[0])
    sd = eng.model.state_dict()
    assert len(sd) == 136
    assert torch.equal(sd["out.weight"].reshape(-1), eng.pview("out.weight").reshape(-1))
    assert int(sd["bottleneck.double_conv.1.num_batches_tracked"]) == 100 + 8
    eng.model.eval()
    y = eng.model(noisy)
    assert y.shape == noisy.shape and torch.isfinite(y).all()


def test_graphed_step_matches_eager():
    """train_step_graphed (one CUDA-graph replay per step) follows the eager step: same losses step by step (the only
    run-to-run freedom is the summation order of the split-K weight-gradient atomics)."""
    noisy, clean = batch(103, (4, 1, 64, 32))
    noisy, clean = noisy.to(DEV), clean.to(DEV)
    a, b = _engine(), _engine()
    for i in range(5):
        la = a.train_step(noisy, clean).cpu()
        lb = b.train_step_graphed(noisy, clean).cpu()
        assert torch.allclose(la, lb, rtol=2e-3), (i, la, lb)
    assert a.step_count == b.step_count == 5 and float(b.step_dev) == 5.0
    assert float((a.P - b.P).abs().max()) <= 1.05e-3          # <= 5 steps x 2 lr: the updates are ~lr * sign(g)


def test_eval_after_graph_replays_uses_the_current_weights():
    """a graph replay updates parameters and BatchNorm running statistics through raw device pointers, so no tensor
    version moves; the module's eval-mode cache (packed bf16 weights, folded BN) must still be refreshed.  train x3 (eager warm-up,
    capture + replay, replay) -> eval -> train x2 (replays) -> eval: each eval forward must equal a FRESH UNet loaded from the
    module's current state_dict, and set_hyperparameters with unchanged values must keep the captured graph."""
    eng = _engine()
    noisy, clean = batch(104, (4, 1, 64, 32))
    noisy, clean = noisy.to(DEV), clean.to(DEV)
    outs = []
    for round_ in range(2):
        for _ in range(3 if round_ == 0 else 2):
            eng.train_step_graphed(noisy, clean)
        graph = eng._graph["graph"]
        eng.set_hyperparameters(lr=eng.lr)                    # unchanged -> no recapture
        assert eng._graph is not None and eng._graph["graph"] is graph
        eng.model.eval()
        y = eng.model(noisy)
        fresh = UNet().eval()
        fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in eng.model.state_dict().items()})
        assert torch.equal(y, fresh(noisy)), f"stale eval cache after graphed steps (round {round_})"
        outs.append(y.clone())
        eng.model.train()
    assert not torch.equal(outs[0], outs[1])
    eng.set_hyperparameters(lr=eng.lr * 0.5)                  # changed -> the baked-in value forces a recapture
    assert eng._graph is None

"""The reference's train_one_epoch body (train.py:61-76), verbatim, on the drop-in UNet / CombinedPerceptualLoss with a stock
torch optimizer -- the autograd integration -- against the fused engine path and the CPU oracle."""
import pytest
import torch

from audiodenoiser_b200 import train as adn_train
from audiodenoiser_b200.checkpoint import seeded_state_dict
from audiodenoiser_b200.loss import CombinedPerceptualLoss
from audiodenoiser_b200.model import UNet
from oracle.make_golden_train import batch
from oracle.train_oracle import TrainOracle

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _loader(n_batches, shape=(4, 1, 64, 32)):
    return [batch(300 + i, shape) for i in range(n_batches)]


def test_reference_loop_runs_unchanged_and_tracks_the_oracle():
    loader = _loader(3)
    net = UNet(); net.load_state_dict(seeded_state_dict(7)); net.to(DEV)
    criterion = CombinedPerceptualLoss()
    optimizer = torch.optim.AdamW(net.parameters(), lr=1e-4)           # train.py:124
    got = adn_train.train_one_epoch(net, loader, criterion, optimizer, DEV)
    for p in net.parameters():
        assert p.grad is not None and p.grad.shape == p.shape
    orc = TrainOracle(seeded_state_dict(7), lr=1e-4)
    ref = sum(float(orc.train_step(a, c)[1][0]) for a, c in loader) / len(loader)
    assert abs(got - ref) <= 5e-2 * abs(ref), (got, ref)
    val = adn_train.validate_one_epoch(net, loader, criterion, DEV)    # eval() forward sees the updated weights / running stats
    assert val == val and val > 0


def test_fused_epoch_matches_autograd_epoch():
    loader = _loader(3)
    a = UNet(); a.load_state_dict(seeded_state_dict(7)); a.to(DEV)
    b = UNet(); b.load_state_dict(seeded_state_dict(7)); b.to(DEV)
    la = adn_train.train_one_epoch(a, loader, CombinedPerceptualLoss(), torch.optim.AdamW(a.parameters(), lr=1e-4), DEV)
    lb = adn_train.train_one_epoch_fused(b, loader, DEV, lr=1e-4)
    assert abs(la - lb) <= 2e-3 * abs(la), (la, lb)
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if k.endswith("num_batches_tracked"):
            assert int(sa[k]) == int(sb[k])
        elif ".double_conv.0.bias" in k or ".double_conv.3.bias" in k:
            continue            # dead parameters: torch applies weight decay to the exact-zero gradient the same way -> still equal
        elif k.endswith(("running_mean", "running_var")):
            # both epochs run the same kernels; they differ in the optimizer (torch AdamW vs the fused one), whose first steps are
            # ~lr * sign(g): one flipped sign in a deep layer moves that layer's batch statistics by a few 1e-3 (measured
            # 3.5e-3 .. 5.0e-3 on the bottleneck across kernel revisions)
            assert float((sa[k] - sb[k]).norm() / sb[k].norm()) <= 1e-2, k
        else:
            assert float((sa[k] - sb[k]).abs().max()) <= 6.5e-4, k      # 3 steps x 2 lr (the update is ~lr * sign(g))


def test_loss_autograd_matches_oracle_gradient():
    from oracle import loss_oracle
    g = torch.Generator().manual_seed(3)
    pred = torch.rand(2, 1, 64, 48, generator=g)
    target = torch.rand(2, 1, 64, 48, generator=g)
    p_ref = pred.clone().requires_grad_()
    tot_ref, s_ref, m_ref, l_ref = loss_oracle.combined_loss(p_ref, target)
    (tot_ref + 0.5 * m_ref).backward()
    p = pred.to(DEV).requires_grad_()
    tot, s, m, l = CombinedPerceptualLoss()(p, target.to(DEV))
    (tot + 0.5 * m).backward()
    assert float((p.grad.cpu() - p_ref.grad).abs().max() / p_ref.grad.abs().max()) < 3e-3

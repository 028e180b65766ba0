"""oracle/train_oracle.py (the CPU restatement of train.py:65-72) against the fixture made from the reference's own
model.py / loss.py / AdamW (oracle/make_golden_train.py).  CPU only."""
import os

import numpy as np
import torch

from audiodenoiser_b200.checkpoint import seeded_state_dict
from oracle.make_golden_train import batch, sample_idx
from oracle.train_oracle import TrainOracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "train_step_small.npz")


def test_train_oracle_matches_reference_two_steps():
    gold = np.load(GOLD)
    torch.manual_seed(0)
    orc = TrainOracle(seeded_state_dict(7), lr=1e-4, max_norm=1.0)
    for step in (1, 2):
        noisy, clean = batch(100 + step)
        out, losses, grads, norm = orc.train_step(noisy, clean)
        assert np.allclose(out.numpy(), gold[f"outputs/{step}"], rtol=1e-4, atol=1e-4)
        assert np.allclose(losses.numpy(), gold[f"losses/{step}"], rtol=1e-4)
        assert abs(float(norm) - float(gold[f"total_norm/{step}"])) <= 1e-3 * float(gold[f"total_norm/{step}"])
        if step == 1:
            for k, g in grads.items():
                ref_norm = float(gold[f"grad_norm/{k}"])
                is_dead_bias = ".double_conv.0.bias" in k or ".double_conv.3.bias" in k      # zero-mean noise under train-mode BN
                if is_dead_bias:
                    assert float(g.norm()) < 1e-4 * float(gold["total_norm/1"])
                    continue
                assert abs(float(g.double().norm()) - ref_norm) <= 2e-3 * ref_norm + 1e-7, k
                s = g.reshape(-1)[sample_idx(g.numel())].numpy()
                assert np.allclose(s, gold[f"grad_samples/{k}"], rtol=5e-3, atol=2e-3 * ref_norm / max(1.0, g.numel() ** 0.5) + 1e-7), k
        sd = orc.state_dict()
        for k, v in sd.items():
            if ".double_conv.0.bias" in k or ".double_conv.3.bias" in k:
                continue                      # driven by the sign of float noise through Adam's normalisation
            flat = v.reshape(-1)
            got = flat[sample_idx(flat.numel())].numpy().astype(np.float64)
            ref = gold[f"state_samples/{step}/{k}"].astype(np.float64)
            assert np.allclose(got, ref, rtol=1e-4, atol=2.5e-4), (k, got, ref)     # one Adam step moves a weight by <= lr = 1e-4

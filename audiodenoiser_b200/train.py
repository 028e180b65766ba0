"""Drop-in for the hot-path functions of the reference's code/train.py: ``train_one_epoch`` (train.py:61-76) and
``validate_one_epoch`` (train.py:78-90), same positional arguments and return value (the epoch's mean total loss).

``train_one_epoch`` is the reference's body verbatim -- it works because ``audiodenoiser_b200.model.UNet`` (train mode) and
``audiodenoiser_b200.loss.CombinedPerceptualLoss`` are autograd nodes backed by the B200 kernels, so ``loss.backward()``,
``clip_grad_norm_`` and any torch optimizer behave as with the reference modules.  ``train_one_epoch_fused`` is the fast path:
the whole step (forward, loss, backward, DDP all-reduce, clip, AdamW) on the engine's kernels, one CUDA-graph replay per batch.
Argument parsing, logging, TensorBoard and checkpoint writing (train.py:20-59,92-151) are CPU glue and out of scope; the
``writer`` / progress-bar arguments are optional here.
"""
from __future__ import annotations

import torch


def train_one_epoch(model, dataloader, criterion, optimizer, device, writer=None, epoch=0):
    """train.py:61-76."""
    model.train()
    total_loss = 0.0
    for noisy, clean in dataloader:
        noisy, clean = noisy.to(device), clean.to(device)            # train.py:65
        optimizer.zero_grad()                                        # :66
        outputs = model(noisy)                                       # :67
        loss, _, _, _ = criterion(outputs, clean)                    # :68
        loss.backward()                                              # :69
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)   # :70
        optimizer.step()                                             # :71
        total_loss += loss.item()                                    # :72
    avg_loss = total_loss / len(dataloader)
    if writer is not None:
        writer.add_scalar("Loss/train", avg_loss, epoch)
    return avg_loss


def validate_one_epoch(model, dataloader, criterion, device, writer=None, epoch=0):
    """train.py:78-90."""
    model.eval()
    total_loss = 0.0
    with torch.no_grad():
        for noisy, clean in dataloader:
            noisy, clean = noisy.to(device), clean.to(device)
            outputs = model(noisy)
            loss, _, _, _ = criterion(outputs, clean)
            total_loss += loss.item()
    avg_loss = total_loss / len(dataloader)
    if writer is not None:
        writer.add_scalar("Loss/validation", avg_loss, epoch)
    return avg_loss


def train_one_epoch_fused(model, dataloader, device, lr=1e-4, writer=None, epoch=0):
    """The same epoch through ``TrainEngine.train_step_graphed``: AdamW(lr) with torch defaults and clip_grad_norm_(1.0) as in
    train.py:70,124, gradients averaged over the ranks of the default process group when one is initialised (DDP)."""
    model.train()
    engine = model.train_engine(device, lr=lr)
    total = torch.zeros((), dtype=torch.float32, device=engine.device)
    for noisy, clean in dataloader:
        losses = engine.train_step_graphed(noisy.to(engine.device, non_blocking=True), clean.to(engine.device, non_blocking=True))
        total += losses[0]
    avg_loss = float(total) / len(dataloader)
    if writer is not None:
        writer.add_scalar("Loss/train", avg_loss, epoch)
    return avg_loss

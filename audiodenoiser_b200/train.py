"""Drop-in for the hot-path functions of the reference's code/train.py: ``train_one_epoch`` (train.py:61-76) and
``validate_one_epoch`` (train.py:78-90), same positional arguments and return value (the epoch's mean total loss).

``train_one_epoch`` is the reference's body verbatim (plus the deferred loader transform for worker batches) -- it works because ``audiodenoiser_b200.model.UNet`` (train mode) and
``audiodenoiser_b200.loss.CombinedPerceptualLoss`` are autograd nodes backed by the B200 kernels, so ``loss.backward()``,
``clip_grad_norm_`` and any torch optimizer behave as with the reference modules.  ``train_one_epoch_fused`` is the fast path:
the whole step (forward, loss, backward, DDP all-reduce, clip, AdamW) on the engine's kernels, one CUDA-graph replay per batch.
Argument parsing, logging, TensorBoard and checkpoint writing (train.py:20-59,92-151) are CPU glue and out of scope; the
``writer`` / progress-bar arguments are optional here.
"""
from __future__ import annotations

import torch

from .data_loader import SpectrogramDataset, finish_on_device


def _deferred_transform(dataloader) -> bool:
    """True when the batches come out of DataLoader WORKER processes over a ``SpectrogramDataset`` (train.py:118-119 uses
    num_workers=4): the workers only crop / pad, and the float16 round trip of data_loader.py:41-42 has to be applied here, on
    the device (see data_loader.py's module docstring)."""
    ds = getattr(dataloader, "dataset", None)
    while hasattr(ds, "dataset"):                 # random_split (train.py:114) wraps the dataset in Subset
        ds = ds.dataset
    return isinstance(ds, SpectrogramDataset) and getattr(dataloader, "num_workers", 0) > 0


def train_one_epoch(model, dataloader, criterion, optimizer, device, writer=None, epoch=0):
    """train.py:61-76."""
    model.train()
    total_loss = 0.0
    finish = _deferred_transform(dataloader)
    for noisy, clean in dataloader:
        noisy, clean = noisy.to(device), clean.to(device)            # train.py:65
        if finish:
            noisy, clean = finish_on_device(noisy), finish_on_device(clean)
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
    finish = _deferred_transform(dataloader)
    with torch.no_grad():
        for noisy, clean in dataloader:
            noisy, clean = noisy.to(device), clean.to(device)
            if finish:
                noisy, clean = finish_on_device(noisy), finish_on_device(clean)
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
    finish = _deferred_transform(dataloader)
    for noisy, clean in dataloader:
        noisy, clean = noisy.to(engine.device, non_blocking=True), clean.to(engine.device, non_blocking=True)
        if finish:
            noisy, clean = finish_on_device(noisy), finish_on_device(clean)
        losses = engine.train_step_graphed(noisy, clean)
        total += losses[0]
    avg_loss = float(total) / len(dataloader)
    if writer is not None:
        writer.add_scalar("Loss/train", avg_loss, epoch)
    return avg_loss

"""Generate tests/golden/noise_small.npz by executing the reference's OWN ``match_audio_length`` / ``add_noise`` definitions.
create_train_dataset.py cannot be imported here (librosa, soundfile, pedalboard are absent), so the two function definitions and the
constants they use are extracted from its source with ``ast`` and executed unmodified in a namespace holding numpy and random.
Run in the build container only.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ast
import os
import random
import sys

import numpy as np

REF = "/root/reference/code/create_train_dataset.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audiodenoiser_b200 import synth  # noqa: E402


def reference_functions():
    tree = ast.parse(open(REF).read())
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("match_audio_length", "add_noise"):
            keep.append(node)
        elif isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id in ("SNR_DB", "SAMPLE_RATE") for t in node.targets):
            keep.append(node)
    ns = {"np": np, "random": random}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF, "exec"), ns)
    return ns


def cases():
    """(name, clean, noise, noise_type, seed)"""
    clip = synth.make_clip(3, "R", return_clean=True)[1]
    out = []
    for k, n in enumerate((16000, 24000, 5000)):
        clean = clip[:n].astype(np.float32)
        rng = np.random.default_rng(50 + k)
        longer = rng.standard_normal(n + 7000).astype(np.float32) * 0.3
        shorter = rng.standard_normal(n // 3 + 11).astype(np.float32) * 0.1
        out += [(f"white_{n}", clean, None, "white", 10 + k), (f"urban_long_{n}", clean, longer, "urban", 20 + k),
                (f"urban_short_{n}", clean, shorter, "urban", 30 + k), (f"urban_none_{n}", clean, None, "urban", 40 + k),
                (f"cancel_{n}", clean, None, "noise_cancellation", 60 + k)]
    out.append(("urban_silent_16000", clip[:16000].astype(np.float32), np.zeros(16000, np.float32), "urban", 70))
    return out


def main():
    ns = reference_functions()
    gold = {}
    for name, clean, noise, nt, seed in cases():
        np.random.seed(seed); random.seed(seed)
        y = ns["add_noise"](clean.copy(), None if noise is None else noise.copy(), nt)
        gold[name] = np.asarray(y)
    path = os.path.join(ROOT, "tests", "golden", "noise_small.npz")
    np.savez_compressed(path, **gold)
    print(path, os.path.getsize(path), "bytes;", {k: (v.dtype.name, v.shape) for k, v in gold.items()})


if __name__ == "__main__":
    main()

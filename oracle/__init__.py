"""CPU oracle for the AudioDenoiser hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / the CPU baseline, never as the
thing shipped.  The product path (``audiodenoiser_b200``) calls hand-written sm_100a
kernels through ``libadn_b200.so`` and raises when that library or a GPU is missing.

Pinning status (see DESIGN.md "Oracle"):
  * ``unet_oracle``  -- pinned: checked against the reference's own ``code/model.py``
    (imported from /root/reference in the build container) through the committed
    fixtures ``tests/golden/unet_*.npz`` made by ``oracle/make_golden.py``.
  * ``loss_oracle``  -- pinned the same way against ``code/loss.py``.
  * ``stft_oracle``  -- PARITY UNPINNED by the reference: the STFT/iSTFT arithmetic
    lives in librosa==0.10.2.post1 (requirements.txt:10), which is neither vendored
    under /root/reference nor installable here, and the reference ships no tests or
    golden vectors.  The restatement follows the published librosa-0.10 algorithm at
    the reference's call sites and is cross-checked against ``torch.stft/istft``
    (float64) and analytic known-answer vectors.
"""

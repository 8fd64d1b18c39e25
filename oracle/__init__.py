"""CPU oracle for the causal-VAD hot path.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker (or as
the timed CPU stand-in for the reference).  The product package never imports it
and fails loudly when its CUDA extension is missing.

What it is: a from-scratch, functional (explicit-parameter, explicit-noise) fp32
restatement of the reference's per-clip forward / loss / optimizer arithmetic in
plain ``torch`` CPU ops.  The reference's arithmetic lives entirely in PyTorch
(third-party, un-pinned: the reference repo ships no requirements file); the torch
2.11.0 CPU kernels are therefore the ground truth both for the reference and here.

Pinning: the reference ships no tests/golden vectors for this path (SURVEY.md
section 4).  The oracle is pinned instead against OUTPUTS OF THE REFERENCE ITSELF run
in the build container: ``tools/make_golden.py`` imports the unmodified scripts
from /root/reference (with noise injection), writes ``tests/golden/*.pt`` and
asserts oracle == reference on every case; ``tests/test_oracle_golden.py`` re-checks
the oracle against those committed fixtures on every run (no reference needed).

Modules (reference file:line each follows is cited per function):
  mb.py    M-B  avenue_training_script2.py:15-205   (3-D CNN + NOTEARS head + 5-term loss)
  mc.py    M-C  minicausal_vad_complete3.py:25-102  (3-D CNN + BN + MLP, BCE)
  ma.py    M-A  causal_anomaly_detection.py:110-586, 649-662
  md.py    M-D  causal_anomaly_detection1.py:124-344, 526-564
  optim.py clip_grad_norm_ / AdamW / Adam restatements (torch.optim semantics)
"""

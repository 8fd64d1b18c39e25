#!/usr/bin/env python
"""A/B timing of the M-A train step (batch 32, one CUDA graph) under switchable code paths, all inside ONE process so a comparison costs
seconds of GPU time instead of one bench.py run per variant.

    python tools/ab_step.py [variant ...]          # default: every variant below; "base" is always measured first

Each variant builds a fresh trainer, captures the step graph with the switches applied, replays it (3 warm-up + 30 timed replays bracketed by
CUDA events) and reports ms/step.  Switches: module attributes of cvad_b200.tc, development modes of the library (cvad_flat_*_mode) and
environment variables that are read at call time.  Development tool; not on the product path.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cvad_b200  # noqa: E402,F401
from cvad_b200 import ops, tc  # noqa: E402
from cvad_b200.ma import CausalAnomalyDetector, MATrainer  # noqa: E402
import synth  # noqa: E402

# name -> list of (kind, key, value);  kind: "tc" module attribute, "lib" library mode function, "env" environment variable
VARIANTS = {
    "base": [],
    "stats_separate": [("tc", "FUSED_STATS", False)],
    "wgrad_direct_atomics": [("tc", "STAGED_WGRAD", False)],
    "wgrad_one_row_per_mma": [("lib", "cvad_flat_wgrad_mode", 0)],
    "no_param_grad_overlap": [("env", "CVAD_PG_OVERLAP", "0")],
}
# switches that exist only on some branches are added when the library exports them
OPTIONAL_LIB_VARIANTS = {"conv_nine_taps": ("cvad_flat_conv_mode", 0), "dgrad_four_launches": ("cvad_flat_dgrad_mode", 0)}
LIB_DEFAULTS = {"cvad_flat_wgrad_mode": 1, "cvad_flat_conv_mode": 1, "cvad_flat_dgrad_mode": 2}


def apply(switches):
    undo = []
    for kind, key, val in switches:
        if kind == "tc":
            undo.append((kind, key, getattr(tc, key)))
            setattr(tc, key, val)
        elif kind == "lib":
            getattr(ops.L(), key)(int(val))
            undo.append((kind, key, LIB_DEFAULTS[key]))
        else:
            undo.append((kind, key, os.environ.get(key)))
            os.environ[key] = val
    return undo


def restore(undo):
    for kind, key, val in undo:
        if kind == "tc":
            setattr(tc, key, val)
        elif kind == "lib":
            getattr(ops.L(), key)(int(val))
        elif val is None:
            os.environ.pop(key, None)
        else:
            os.environ[key] = val


def measure(x, y, dev, replays=30):
    torch.manual_seed(1234)
    tr = MATrainer(CausalAnomalyDetector(), dev, precision="bf16")
    tr.model.train()
    gs = tr.graphed_train_step(x, y)
    xs, ys = gs.static_inputs
    for _ in range(3):
        gs(xs, ys)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        out = gs(xs, ys)
    e1.record()
    torch.cuda.synchronize()
    loss = float(out[0][0])
    del gs, tr
    torch.cuda.empty_cache()
    return e0.elapsed_time(e1) / replays, loss


def main():
    dev = torch.device("cuda:0")
    for name, (fn, val) in OPTIONAL_LIB_VARIANTS.items():
        if hasattr(ops.L(), fn):
            VARIANTS[name] = [("lib", fn, val)]
    names = [n for n in sys.argv[1:] if n in VARIANTS] or list(VARIANTS)
    if "base" in names:
        names.remove("base")
    x = synth.ma_clips(32, 16, 240, 360, 1234, wide=True).to(dev)
    y = (torch.rand(32, generator=synth.gen(1243)) < 0.3).long().to(dev)
    base, loss = measure(x, y, dev)
    print(f"{'base':28s} {base:7.3f} ms/step   loss {loss:.5f}")
    for n in names:
        undo = apply(VARIANTS[n])
        try:
            ms, loss = measure(x, y, dev)
        finally:
            restore(undo)
        print(f"{n:28s} {ms:7.3f} ms/step   loss {loss:.5f}   ({ms - base:+.3f} ms vs base)")


if __name__ == "__main__":
    main()

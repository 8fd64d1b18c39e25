"""Oracle (TEST INFRASTRUCTURE): whole training / inference steps of the reference restated on CPU tensors.

Used by tests (trajectory checks), by ``bench.py``'s ``cpu_baseline`` leg and by ``bench.py --impl reference`` as the
timed CPU stand-in for the reference (``/root/reference`` does not exist on the GPU box, so ``kind`` is "port").
Step structure follows cad:637-693 (M-A), s2:217-238 (M-B) and mc3:256-311 (M-C).
"""
from __future__ import annotations

import torch

from . import ma as o_ma
from . import mb as o_mb
from . import mc as o_mc
from . import optim as o_opt


class OracleAdam:
    """Per-tensor Adam/AdamW state around oracle.optim, with clip_grad_norm_ and the skip-None-grad rule."""

    def __init__(self, names, lr, weight_decay, decoupled, max_norm, clip_threshold=None):
        self.names, self.lr, self.wd, self.decoupled = list(names), lr, weight_decay, decoupled
        self.max_norm, self.clip_threshold = max_norm, clip_threshold
        self.m, self.v, self.t = {}, {}, {}

    def step(self, P, grads):
        live = [k for k in self.names if grads.get(k) is not None]
        norm, clipped = o_opt.clip_grad_norm([grads[k] for k in live], self.max_norm)
        if self.clip_threshold is not None and float(norm) <= self.clip_threshold:
            clipped = [grads[k] for k in live]
        fn = o_opt.adamw_step if self.decoupled else o_opt.adam_step
        for k, g in zip(live, clipped):
            if k not in self.m:
                self.m[k], self.v[k], self.t[k] = torch.zeros_like(P[k]), torch.zeros_like(P[k]), 0
            self.t[k] += 1
            P[k], self.m[k], self.v[k] = fn(P[k], g, self.m[k], self.v[k], self.t[k], self.lr, weight_decay=self.wd)
        return float(norm)


def _leaf(P, trainable):
    return {k: (v.detach().clone().requires_grad_(True) if k in trainable else v) for k, v in P.items()}


def ma_trainable(P):
    return [k for k, v in P.items() if v.is_floating_point() and "running" not in k and "backbone.conv1" not in k
            and "backbone.bn1" not in k]


def ma_train_step(P, opt, x, labels, eps, keep):
    """One iteration of cad:641-690 (fp32 branch).  Mutates P (parameters and BN buffers); returns (loss, comps)."""
    names = ma_trainable(P)
    Pg = _leaf(P, set(names))
    ns = {}
    out = o_ma.ma_forward(Pg, x, eps, True, keep, ns)
    loss, comps = o_ma.ma_loss(out, labels)
    loss.backward()
    grads = {}
    for k in names:
        g = Pg[k].grad
        grads[k] = None if g is None or (float(g.abs().max()) == 0.0 and ("detector" in k or "structure_learner" in k)) else g
    opt.step(P, grads)
    P.update(ns)
    return float(loss), comps


def mb_train_step(P, opt, x, pseudo, keep_feat, keep_graph):
    """One iteration of s2:221-238."""
    names = list(P.keys())
    Pg = _leaf(P, set(names))
    s, a, _ = o_mb.mb_forward(Pg, x, True, keep_feat, keep_graph)
    loss, comps = o_mb.mb_loss(s, a, pseudo)
    loss.backward()
    opt.step(P, {k: Pg[k].grad for k in names})
    return float(loss), comps


def mc_infer(P, x):
    with torch.no_grad():
        return o_mc.mc_forward(P, x)

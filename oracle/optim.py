"""Oracle (TEST INFRASTRUCTURE): restatement of the optimizer-step conventions on the hot path.

torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW / Adam single-tensor semantics as used at
s2:115-119, 236-238 (AdamW lr 5e-4 wd 1e-3, clip 0.5), cad:615-618, 665-667 (AdamW lr 3e-4 wd 1e-5,
clip 1.0) and mc3:229-234, 298-311 (Adam with L2 weight decay 1e-5; clip to 1.0 only when the
norm exceeds 10).  Tensors whose grad is None are skipped entirely (no decay, no state update).
"""
from __future__ import annotations

import math

import torch


def clip_grad_norm(grads, max_norm: float):
    """Returns (total_norm, clipped grads).  coef = max_norm / (norm + 1e-6), clamped to 1."""
    gs = [g for g in grads if g is not None]
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in gs)).float() if gs else torch.zeros(())
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, [None if g is None else g * coef for g in grads]


def adamw_step(p, g, m, v, step: int, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2):
    """One decoupled-weight-decay Adam step (torch.optim.AdamW, amsgrad=False).  step is 1-based."""
    p = p * (1 - lr * weight_decay)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def adam_step(p, g, m, v, step: int, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam: L2 weight decay is folded into the gradient."""
    g = g + weight_decay * p
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v

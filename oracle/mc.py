"""Oracle (TEST INFRASTRUCTURE) for M-C: SimpleVideoAnomalyDetector.

Follows minicausal_vad_complete3.py (mc3): features mc3:36-57, classifier mc3:60-69,
forward mc3:90-102, BCE mc3:240/287.  ``P`` holds parameters and BN buffers under the
reference's state_dict names (features.{0,1,4,5,8,9}.*, classifier.{1,4,6}.*).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOM = 0.1
POOLS = ((1, 2, 2), (2, 2, 2), (2, 2, 2))   # mc3:41, 47, 53
CONV_IDX = (0, 4, 8)
BN_IDX = (1, 5, 9)


def _bn(h, P, idx, train, new_stats):
    w, b = P[f"features.{idx}.weight"], P[f"features.{idx}.bias"]
    if train:
        dims = (0, 2, 3, 4)
        mean = h.mean(dim=dims)
        var = h.var(dim=dims, unbiased=False)
        n = h.numel() // h.shape[1]
        if new_stats is not None:
            new_stats[f"features.{idx}.running_mean"] = (1 - BN_MOM) * P[f"features.{idx}.running_mean"] + BN_MOM * mean.detach()
            new_stats[f"features.{idx}.running_var"] = (1 - BN_MOM) * P[f"features.{idx}.running_var"] + BN_MOM * var.detach() * n / max(n - 1, 1)
            new_stats[f"features.{idx}.num_batches_tracked"] = P[f"features.{idx}.num_batches_tracked"] + 1
    else:
        mean, var = P[f"features.{idx}.running_mean"], P[f"features.{idx}.running_var"]
    sh = (1, -1, 1, 1, 1)
    return (h - mean.view(sh)) / torch.sqrt(var.view(sh) + BN_EPS) * w.view(sh) + b.view(sh)


def mc_forward(P: dict, x: torch.Tensor, train: bool = False, keep0=None, keep1=None, new_stats: dict | None = None):
    """x (B,1,T,H,W) -> scores (B,1).  keep0 (B,32) for Dropout(0.5) mc3:61, keep1 (B,16) for Dropout(0.3) mc3:64."""
    if x.dim() != 5:
        raise ValueError(f"Expected 5D tensor (B,C,T,H,W), got {tuple(x.shape)}")   # mc3:92-93
    h = x
    for ci, bi, pool in zip(CONV_IDX, BN_IDX, POOLS):
        h = F.conv3d(h, P[f"features.{ci}.weight"], P[f"features.{ci}.bias"], stride=1, padding=1)
        h = F.relu(_bn(h, P, bi, train, new_stats))
        h = F.max_pool3d(h, kernel_size=pool, stride=pool)
    f = h.mean(dim=(2, 3, 4))                               # AdaptiveAvgPool3d(1), mc3:56
    if train:
        f = f * keep0 * 2.0
    h = F.relu(F.linear(f, P["classifier.1.weight"], P["classifier.1.bias"]))
    if train:
        h = h * keep1 * (1.0 / 0.7)
    h = F.relu(F.linear(h, P["classifier.4.weight"], P["classifier.4.bias"]))
    return torch.sigmoid(F.linear(h, P["classifier.6.weight"], P["classifier.6.bias"]))


def bce(scores: torch.Tensor, targets: torch.Tensor):
    """nn.BCELoss (mean), log terms clamped at -100.  mc3:240, 287."""
    s = scores.reshape(-1)
    return -(targets * torch.clamp(torch.log(s), min=-100.0)
             + (1 - targets) * torch.clamp(torch.log(1 - s), min=-100.0)).mean()

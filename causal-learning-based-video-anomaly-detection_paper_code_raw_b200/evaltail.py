"""Evaluation tail on the device: thresholds, pseudo-labels, ROC-AUC, evaluation metrics and score smoothing computed
where the scores already are, instead of copying every score and every (16,16) graph to the host for numpy / sklearn.

Reference call sites: ``np.percentile`` s1:60, cad1:609, cad1:709; ``roc_auc_score`` mc3:388, cad:1233-1248; the eight metrics of
s2:286-295 (incl. ``len(np.unique(graphs, axis=0))``); ``np.convolve(scores, ones(w)/w, 'valid')`` cad:1085-1087, vad:833-835.
Every function takes CUDA tensors and returns CUDA tensors (no synchronisation); ``.item()`` / ``.tolist()`` on the result is the one
read-back of a handful of scalars.
"""
from __future__ import annotations

import torch

from . import ops
from .ops import _call, _ptr, _st

METRIC_KEYS = ("mean_score", "std_score", "min_score", "max_score", "score_range", "avg_edges", "avg_sparsity", "unique_graphs")


def _scores(x):
    ops._cuda(x)
    return x.detach().reshape(-1).float().contiguous()


def _workspace(n, device):
    nbytes = int(ops.L().cvad_eval_workspace_bytes(int(n)))
    return torch.empty(nbytes, device=device, dtype=torch.uint8)


def sort_scores(scores):
    """Ascending (values, order) like ``np.sort`` / ``np.argsort`` (NaNs last; ties in index order)."""
    x = _scores(scores)
    n = x.numel()
    out = torch.empty_like(x)
    order = torch.empty(n, device=x.device, dtype=torch.int32)
    if n:
        _call("cvad_sort_scores_f32", _ptr(x), n, _ptr(_workspace(n, x.device)), _ptr(out), _ptr(order), _st())
    return out, order


def percentile(scores, q: float):
    """``np.percentile(scores, q)`` (float32 scores, method 'linear') as a 1-element CUDA tensor, bit-identical to numpy."""
    x = _scores(scores)
    srt, _ = sort_scores(x)
    out = torch.zeros(1, device=x.device, dtype=torch.float32)
    _call("cvad_percentile_sorted_f32", _ptr(srt), srt.numel(), float(q), _ptr(out), _st())
    return out


def threshold_labels(scores, threshold):
    """``(scores > threshold).astype(float)`` with the threshold read from device memory (s1:61)."""
    x = _scores(scores)
    labels = torch.empty_like(x)
    _call("cvad_threshold_labels_f32", _ptr(x), x.numel(), _ptr(threshold), _ptr(labels), _st())
    return labels


def percentile_labels(scores, q: float = 95.0):
    """s1:57-61 in one go: (threshold (1,), pseudo-labels (N,)), nothing leaves the device."""
    thr = percentile(scores, q)
    return thr, threshold_labels(scores, thr)


def roc_auc(scores, targets):
    """``sklearn.metrics.roc_auc_score(targets, scores)`` as a 1-element fp64 CUDA tensor (0.0 when only one class is present)."""
    x = _scores(scores)
    t = _scores(targets)
    if t.numel() != x.numel():
        raise ValueError("scores and targets differ in length")
    auc = torch.zeros(1, device=x.device, dtype=torch.float64)
    _call("cvad_roc_auc_f32", _ptr(x), _ptr(t), x.numel(), _ptr(_workspace(max(x.numel(), 1), x.device)), _ptr(auc), _st())
    return auc


def mb_eval_metrics(scores, graphs, edge_threshold: float = 0.1):
    """The eight entries of s2:286-295 as an (8,) fp64 CUDA tensor in METRIC_KEYS order."""
    x = _scores(scores)
    ops._cuda(graphs)
    n = x.numel()
    g = graphs.detach().reshape(n, -1).float().contiguous()
    out = torch.zeros(8, device=x.device, dtype=torch.float64)
    _call("cvad_mb_eval_metrics_f32", _ptr(x), _ptr(g), n, g.shape[1] if n else 256, float(edge_threshold), _ptr(_workspace(max(n, 1), x.device)),
          _ptr(out), _st())
    return out


def moving_average(scores, window: int):
    """``np.convolve(scores, np.ones(window) / window, mode='valid')`` (fp64, length N - window + 1; empty when N < window)."""
    x = _scores(scores)
    n = x.numel()
    out = torch.empty(max(n - window + 1, 0), device=x.device, dtype=torch.float64)
    if out.numel():
        _call("cvad_moving_average_f32", _ptr(x), n, int(window), _ptr(out), _st())
    return out

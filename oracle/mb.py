"""Oracle (TEST INFRASTRUCTURE) for M-B: the checkpointed 3-D CNN + NOTEARS-style head.

Follows avenue_training_script2.py (s2):
  forward  s2:27-35 (feature extractor), s2:50-60 (causal discovery), s2:91-101 (detector)
  loss     s2:135-205 (focal BCE on pseudo-labels, acyclicity, sparsity, consistency, structure)
Parameters are passed as a flat dict with the checkpoint's key names (s2:437-443).
All randomness is explicit: dropout keep-masks and the pseudo-label draw are arguments.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

P_FEAT_DROP = 0.3   # s2:25
P_GRAPH_DROP = 0.3  # s2:80


def mb_forward(P: dict, x: torch.Tensor, train: bool = False,
               keep_feat: torch.Tensor | None = None, keep_graph: torch.Tensor | None = None):
    """x (B,3,T,H,W) fp32 -> (scores (B,1), adj (B,16,16), feat (B,16)).  s2:91-101."""
    fe = "feature_extractor."
    h = F.relu(F.conv3d(x, P[fe + "conv3d_1.weight"], P[fe + "conv3d_1.bias"], stride=(1, 2, 2), padding=1))
    h = F.relu(F.conv3d(h, P[fe + "conv3d_2.weight"], P[fe + "conv3d_2.bias"], stride=2, padding=1))
    h = F.relu(F.conv3d(h, P[fe + "conv3d_3.weight"], P[fe + "conv3d_3.bias"], stride=2, padding=1))
    h = F.adaptive_avg_pool3d(h, (4, 4, 4))                      # s2:23, 32
    feat = F.linear(h.flatten(1), P[fe + "fc.weight"], P[fe + "fc.bias"])
    if train:                                                    # dropout on the feature itself, s2:34
        feat = feat * keep_feat * (1.0 / (1.0 - P_FEAT_DROP))
    B = feat.shape[0]
    a = F.relu(F.linear(feat, P["causal_discovery.causal_net.0.weight"], P["causal_discovery.causal_net.0.bias"]))
    a = torch.sigmoid(F.linear(a, P["causal_discovery.causal_net.2.weight"], P["causal_discovery.causal_net.2.bias"]))
    adj = a.view(B, 16, 16) * (1.0 - torch.eye(16))              # s2:57-58
    g = F.relu(F.linear(adj.reshape(B, -1), P["graph_encoder.0.weight"], P["graph_encoder.0.bias"]))
    if train:
        g = g * keep_graph * (1.0 / (1.0 - P_GRAPH_DROP))
    g = F.linear(g, P["graph_encoder.3.weight"], P["graph_encoder.3.bias"])
    c = torch.cat([feat, g], dim=1)                              # s2:98
    s = F.relu(F.linear(c, P["anomaly_predictor.0.weight"], P["anomaly_predictor.0.bias"]))
    s = torch.sigmoid(F.linear(s, P["anomaly_predictor.2.weight"], P["anomaly_predictor.2.bias"]))
    return s, adj, feat


def mb_loss(scores: torch.Tensor, adj: torch.Tensor, pseudo: torch.Tensor,
            anomaly_weight=1.0, causal_weight=0.01, sparsity_weight=0.001, consistency_weight=0.01):
    """s2:135-205 with the pseudo-label vector ``pseudo`` (B,) in {0,1} supplied by the caller
    (the reference draws it as ``rand_like(targets) > 0.95``, s2:141; true labels are unused).
    Returns (total, dict of the 7 reported components as python floats)."""
    s = scores.reshape(-1)
    # focal BCE, s2:144-149 (F.binary_cross_entropy clamps each log term at -100)
    ce = -(pseudo * torch.clamp(torch.log(s), min=-100.0) + (1 - pseudo) * torch.clamp(torch.log(1 - s), min=-100.0))
    pt = torch.exp(-ce)
    anomaly = (0.25 * (1 - pt) ** 2 * ce).mean()
    # acyclicity, s2:152-153
    abar = adj.mean(dim=0)
    acyc = torch.trace(abar @ abar)
    # sparsity (no gradient), s2:156-158
    cur_sparsity = (adj > 0.1).float().mean()
    spars = (cur_sparsity - 0.3).abs()
    # consistency over "normal" clips, s2:161-177 (all unordered pairs i<j)
    normal = adj[pseudo == 0]
    n = normal.shape[0]
    if n > 1:
        d = (normal[:, None] - normal[None, :]).abs().mean(dim=(2, 3))     # (n,n)
        iu = torch.triu_indices(n, n, offset=1)
        cons = (d[iu[0], iu[1]].mean() - 0.1).abs()
    else:
        cons = torch.zeros(())
    # structure hinge on the batch-wide edge count (no gradient), s2:180-189
    edges = (adj > 0.1).sum()
    if edges < 10:
        struct = (10 - edges) * 0.01
    elif edges > 40:
        struct = (edges - 40) * 0.01
    else:
        struct = torch.zeros(())
    total = (anomaly_weight * anomaly + causal_weight * acyc + sparsity_weight * spars
             + consistency_weight * cons + 0.01 * struct)
    comps = {
        "anomaly_loss": float(anomaly.detach()), "acyclicity_loss": float(acyc.detach()), "sparsity_loss": float(spars),
        "consistency_loss": float(cons.detach()), "structure_loss": float(struct), "edge_count": float(edges),
        "sparsity_ratio": float(cur_sparsity),
    }
    return total, comps


def mb_eval_metrics(pred, graphs):
    """s2:286-295 on numpy arrays (N,), (N,16,16)."""
    import numpy as np
    e = np.sum(graphs > 0.1, axis=(1, 2))
    return {
        "mean_score": float(np.mean(pred)), "std_score": float(np.std(pred)),
        "min_score": float(np.min(pred)), "max_score": float(np.max(pred)),
        "score_range": float(np.max(pred) - np.min(pred)),
        "avg_edges": float(np.mean(e)), "avg_sparsity": float(np.mean(e / 256)),
        "unique_graphs": len(np.unique(graphs.reshape(len(graphs), -1), axis=0)),
    }

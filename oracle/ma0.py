"""Oracle (TEST INFRASTRUCTURE) for M-A0: the older detector / tracker / GRU / VAE / graph model.

Follows video_anomaly_detection.py (vad): backbone vad:67-115 (same as cad's), detector vad:117-165,
tracker vad:167-215, trajectory encoder vad:217-252, factor extractor vad:254-296, structure learner
vad:298-344, dynamics vad:346-373, scorer vad:375-403, model vad:405-454, 2-term loss vad:516-531.

Dense, masked, batched restatement (tracks padded to 5 slots per clip with a per-clip count, like
oracle/ma.py); tools/make_golden.py proves it equal to the reference's ragged Python lists, including
frames where 0, 1, 2 or 3 anchors pass the confidence threshold.  The only noise is ``eps`` (B,5,6)
for the reparameterisation (drawn in eval mode too, vad:270-273).  Only tests/, smoke() and bench.py's
CPU legs may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .ma import MAXDET, NF, _mlp, backbone, gru_last

NA = 3


def detect0(P, feats):
    """vad:127-165.  Boxes (B,T,5,4): the anchors with confidence > 0.5 in descending-confidence order (raw regressions), an all-zero
    dummy row when none passes; counts (B,T) >= 1."""
    B, T, _ = feats.shape
    bbox = F.linear(feats, P["detector.bbox_head.weight"], P["detector.bbox_head.bias"]).view(B, T, NA, 4)
    conf = torch.sigmoid(F.linear(feats, P["detector.conf_head.weight"], P["detector.conf_head.bias"])).detach()
    out = torch.zeros(B, T, MAXDET, 4)
    cnt = torch.ones(B, T, dtype=torch.int64)
    real = torch.zeros(B, T, dtype=torch.int64)        # anchors that passed (0 = the dummy box stands in)
    rows = []
    for b in range(B):
        for t in range(T):
            order = sorted(range(NA), key=lambda a: (-float(conf[b, t, a]), a))
            keep = [a for a in order if float(conf[b, t, a]) > 0.5]
            if keep:
                rows.append((b, t, keep))
                cnt[b, t] = len(keep)
                real[b, t] = len(keep)
    if rows:                                   # one differentiable scatter instead of in-place writes
        pieces = []
        for b, t, keep in rows:
            sel = bbox[b, t, keep]
            pad = torch.zeros(MAXDET - len(keep), 4)
            pieces.append((b * T + t, torch.cat([sel, pad], dim=0)))
        flat = [torch.zeros(MAXDET, 4) for _ in range(B * T)]
        for i, v in pieces:
            flat[i] = v
        out = torch.stack(flat).view(B, T, MAXDET, 4)
    return out, cnt, real


def ma0_forward(P: dict, x: torch.Tensor, eps: torch.Tensor, train: bool = False, new_stats: dict | None = None):
    """x (B,T,1,H,W) -> dict mirroring vad:448-454 with dense tensors: anomaly_scores (B,), causal_factors (B,5,6),
    adjacency_matrices (B,6,6), kl_losses (B,), detections (B,T,5,4) + det_counts (B,T), n_tracks (B,)."""
    feats = backbone(P, x, train, new_stats)
    B, T, _ = feats.shape
    box, cnt, real = detect0(P, feats)
    ntr = cnt.max(dim=1).values                                              # tracks per clip = longest frame list, vad:199
    kidx = torch.arange(MAXDET)
    row_ok = (kidx.view(1, 1, -1) < cnt.unsqueeze(-1)).float().unsqueeze(-1)
    reid = _mlp(P, "tracker.reid_net", (0, 2, 4), box, ("relu", "relu", None))
    traj = torch.cat([box, reid], dim=-1) * row_ok                           # zero padding rows, vad:203-205
    trk_ok = (kidx.view(1, -1) < ntr.unsqueeze(-1)).float()
    hT = gru_last(P, traj.permute(0, 2, 1, 3).reshape(B * MAXDET, T, -1))
    enc = F.linear(hT, P["traj_encoder.encoder.weight"], P["traj_encoder.encoder.bias"]).view(B, MAXDET, -1)
    h = _mlp(P, "causal_extractor.encoder", (0, 2), enc, ("relu", "relu"))
    mu = F.linear(h, P["causal_extractor.mu_head.weight"], P["causal_extractor.mu_head.bias"])
    lv = F.linear(h, P["causal_extractor.logvar_head.weight"], P["causal_extractor.logvar_head.bias"])
    z = mu + eps * torch.exp(0.5 * lv)                                       # vad:270-273
    klrow = -0.5 * (1 + lv - mu.pow(2) - lv.exp()).sum(dim=-1)               # vad:286
    kl = (klrow * trk_ok).sum(dim=1) / ntr
    node = F.linear(z, P["structure_learner.node_encoder.weight"], P["structure_learner.node_encoder.bias"])
    pair = torch.cat([node.unsqueeze(2).expand(B, MAXDET, MAXDET, -1), node.unsqueeze(1).expand(B, MAXDET, MAXDET, -1)], dim=-1)
    e = _mlp(P, "structure_learner.edge_predictor", (0, 2), pair, ("relu", "sigmoid")).squeeze(-1)
    emask = trk_ok.unsqueeze(2) * trk_ok.unsqueeze(1) * (1 - torch.eye(MAXDET))
    adj = torch.zeros(B, NF, NF)
    adj[:, :MAXDET, :MAXDET] = e * emask                                     # vad:324-335
    structured = torch.einsum("bij,bkj->bki", adj, z)                        # vad:364
    pred = _mlp(P, "dynamics_predictor.dynamics_net", (0, 2, 4), structured, ("relu", "relu", None))
    rows = torch.cat([z, pred, (z - pred).abs()], dim=-1)                    # vad:392-395, per track
    s = _mlp(P, "anomaly_scorer.score_net", (0, 2, 4), rows, ("relu", "relu", "sigmoid")).squeeze(-1)
    scores = (s * trk_ok).sum(dim=1) / ntr                                   # vad:397
    return {"anomaly_scores": scores, "causal_factors": z, "adjacency_matrices": adj, "kl_losses": kl, "detections": box,
            "det_counts": cnt, "det_real": real, "n_tracks": ntr, "features": feats}


def ma0_loss(out: dict, labels: torch.Tensor):
    """vad:516-531 (the standard-precision branch): MSE + 0.001 * mean of the finite KL terms."""
    mse = F.mse_loss(out["anomaly_scores"], labels.float())
    kl = out["kl_losses"]
    fin = torch.isfinite(kl)
    n = int(fin.sum())
    klm = torch.where(fin, kl, torch.zeros_like(kl)).sum() / n if n else torch.zeros(())
    total = mse + 0.001 * klm
    return total, {"anomaly": float(mse.detach()), "kl": float(klm.detach())}

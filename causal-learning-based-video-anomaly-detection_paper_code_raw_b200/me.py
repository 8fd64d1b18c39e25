"""M-E: the visualiser's stand-in model and its sliding-window clip scoring (avenue_training_script_bbox.py), on cvad_b200 kernels.

Mirrors ``CausalAnomalyDetector`` bbox:51-101 (the reference's own docstring calls it a placeholder; it is what
``AnomalyVisualizer`` instantiates, bbox:128), ``predict_anomaly_for_clip`` bbox:328-357 and the window generator of
``extract_anomalous_frames`` bbox:392-430 (8-frame windows, stride 4, one batch-1 forward + three device->host copies per
window in the reference).  Here all windows of a video are scored in batches with one read-back; person detection, drawing
and the HTML report (bbox:157-326, 432-659) are out of scope (third-party detectors, OpenCV drawing).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .noise import DeviceNoise
from .ops import ACT_NONE, ACT_RELU, ACT_SIGMOID


class CausalAnomalyDetector(nn.Module):
    """bbox:51-101.  forward((B,3,T,H,W)) -> (anomaly_score.squeeze(), causal_adj (B,16,16), features (B,1024))."""

    def __init__(self, input_channels=3, hidden_dim=64, num_frames=8):
        super().__init__()
        self.num_frames = num_frames
        self.encoder = nn.Sequential(nn.Conv3d(input_channels, 32, kernel_size=3, stride=1, padding=1), nn.ReLU(), nn.MaxPool3d(2),
                                     nn.Conv3d(32, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(), nn.AdaptiveAvgPool3d((1, 4, 4)))
        self.feature_dim = 64 * 16
        self.causal_net = nn.Sequential(nn.Linear(self.feature_dim, 256), nn.ReLU(), nn.Linear(256, 16 * 16))
        self.classifier = nn.Sequential(nn.Linear(self.feature_dim, 128), nn.ReLU(), nn.Dropout(0.3), nn.Linear(128, 1), nn.Sigmoid())
        self.noise = DeviceNoise()

    def forward(self, x):
        B = x.size(0)
        e = self.encoder
        h = ops.conv_act(x, e[0].weight, e[0].bias, 1, 1, ACT_RELU)
        h = ops.maxpool(h, 2)
        h = ops.conv_act(h, e[3].weight, e[3].bias, 1, 1, ACT_RELU)
        features = ops.adaptive_avgpool(h, (1, 4, 4)).reshape(B, -1)
        c = self.causal_net
        adj = ops.linear_act(ops.linear_act(features, c[0].weight, c[0].bias, ACT_RELU), c[2].weight, c[2].bias, ACT_SIGMOID).view(B, 16, 16)
        k = self.classifier
        keep = self.noise.keep_mask("cls", (B, 128), 0.3, x.device) if self.training else None
        s = ops.linear_act(ops.linear_act(features, k[0].weight, k[0].bias, ACT_RELU, keep, 0.3), k[3].weight, k[3].bias, ACT_SIGMOID)
        return s.squeeze(), adj, features


def load_checkpoint(model, path, device):
    """bbox:128-155: accepts {'model_state_dict'}, {'state_dict'} or a bare state_dict.  Unlike the reference (which swallows the
    mismatch and silently keeps random weights, bbox:150-155) a checkpoint that does not fit raises."""
    ck = torch.load(path, map_location=device, weights_only=False)
    sd = ck.get("model_state_dict", ck.get("state_dict", ck)) if isinstance(ck, dict) else ck
    model.load_state_dict(sd, strict=True)
    return model


@torch.no_grad()
def predict_anomaly_for_clip(model, video_clip, device="cuda"):
    """bbox:328-357: one clip (3,T,H,W) or (1,3,T,H,W), numpy or tensor -> (score float, adjacency (16,16), features)."""
    t = torch.from_numpy(video_clip).float() if isinstance(video_clip, np.ndarray) else video_clip.float()
    if t.dim() == 4:
        t = t.unsqueeze(0)
    s, adj, feat = model(t.to(device))
    return float(s.reshape(-1)[0]), adj[0].cpu().numpy(), feat[0].cpu().numpy()


@torch.no_grad()
def score_windows(model, frames, window=8, stride=4, batch=256, device="cuda"):
    """All sliding windows of one video in batches (bbox:392-415 scores them one by one).

    frames: (F,3,H,W) float in [0,1] (what bbox:397-411 builds per window: resize, BGR->RGB, /255).  Windows start at
    ``range(0, F - window, stride)`` exactly as the reference.  Returns (starts (n,), scores (n,), adj (n,16,16), features)."""
    model.eval()
    dev = torch.device(device)
    F_ = frames.shape[0]
    starts = list(range(0, F_ - window, stride))
    if not starts:
        z = torch.zeros(0)
        return np.zeros(0, dtype=np.int64), z.numpy(), np.zeros((0, 16, 16), np.float32), np.zeros((0, model.feature_dim), np.float32)
    scores, adjs, feats = score_windows_device(model, frames.to(dev, non_blocking=True), window, stride, batch)
    return (np.asarray(starts), scores.cpu().numpy(), adjs.cpu().numpy(), feats.cpu().numpy())


@torch.no_grad()
def score_windows_device(model, frames_dev, window=8, stride=4, batch=256):
    """The device part of ``score_windows``: frames (F,3,H,W) already on the GPU -> (scores (n,), adj (n,16,16), features) device tensors,
    no host synchronisation."""
    dev = frames_dev.device
    F_ = frames_dev.shape[0]
    starts = list(range(0, F_ - window, stride))
    fr = frames_dev.float()
    win = fr.unfold(0, window, 1)                         # (F-window+1, 3, H, W, window) view
    idx = torch.tensor(starts, device=dev)
    scores, adjs, feats = [], [], []
    for i in range(0, len(starts), batch):
        clips = win[idx[i:i + batch]].permute(0, 1, 4, 2, 3).contiguous()       # (b, 3, window, H, W)
        s, a, f = model(clips)
        scores.append(s.reshape(-1))
        adjs.append(a)
        feats.append(f)
    return torch.cat(scores), torch.cat(adjs), torch.cat(feats)


def extract_anomalous_windows(model, frames, video_id="video", threshold=0.3, window=8, stride=4, device="cuda"):
    """The records bbox:416-425 collects for windows whose score exceeds the threshold (frame paths replaced by indices)."""
    starts, scores, adjs, feats = score_windows(model, frames, window, stride, device=device)
    out = []
    for s0, sc, a, f in zip(starts, scores, adjs, feats):
        if sc > threshold:
            out.append({"video_id": video_id, "start_frame": int(s0), "end_frame": int(s0 + window), "anomaly_score": float(sc),
                        "causal_graph": a, "features": f})
    return out

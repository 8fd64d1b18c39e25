"""M-A0: the older detector / tracker / GRU / VAE / causal-graph model of video_anomaly_detection.py ("vad"), on cvad_b200 kernels,
and streaming sliding-window inference for the per-frame (2-D backbone) models.

Mirrors vad: ``ResNetBackbone`` vad:67-115 (identical to cad's), ``PedestrianDetector`` vad:117-165, ``TrajectoryTracker``
vad:167-215, ``TrajectoryEncoder`` vad:217-252, ``CausalFactorExtractor`` vad:254-296, ``CausalStructureLearner`` vad:298-344,
``DynamicsPredictor`` vad:346-373, ``AnomalyScorer`` vad:375-403, ``CausalAnomalyDetector`` vad:405-454,
``apply_memory_efficient_training`` vad:456-472, ``train_model`` vad:474-637, ``test_model`` vad:639-657.  Module / parameter names
(hence ``state_dict`` keys) and the 5-key output dict are the reference's.  What differs from M-A (ma.py): the detector keeps the
anchors whose confidence exceeds 0.5 in top-k order (raw box regressions, no squashing, an all-zero dummy box otherwise), one scorer
MLP runs on every track row and is averaged over the clip's tracks, there is no direct classifier, and the loss has two terms.
Everything else (tracker, GRU, VAE head, structure learner, dynamics) is shared with ma.py, as are the dense masked batches
(5 track slots per clip + a count) that replace the ragged Python lists.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ma_ops, ops
from .arena import FusedAdam
from .ma import (SLOT_DETECTOR, SLOT_NEVER, SLOT_STRUCTURE, CausalFactorExtractor, CausalStructureLearner, DynamicsPredictor, ResNetBackbone,
                 TrajectoryEncoder, TrajectoryTracker, _LazyList, _mlp, apply_memory_efficient_training)
from .noise import DeviceNoise
from .ops import ACT_NONE, ACT_RELU, ACT_SIGMOID, _call, _cuda, _f32c, _ptr, _st

MAXDET = ma_ops.MAXDET


# ------------------------------------------------------------------------------------------------ autograd glue (csrc/ma0_tail.cu)
class _DetTopkDecode(torch.autograd.Function):
    """bbox (B,T,3,4), conf logits (B,T,3) -> box (B,T,5,4); cnt (B,T) int32 and src (B,T,5) int32 are non-differentiable."""

    @staticmethod
    def forward(ctx, bbox, conf_logit, flag):
        _cuda(bbox, conf_logit)
        bbox, conf_logit = _f32c(bbox), _f32c(conf_logit)
        B, T = bbox.shape[:2]
        box = torch.empty((B, T, MAXDET, 4), device=bbox.device, dtype=torch.float32)
        cnt = torch.empty((B, T), device=bbox.device, dtype=torch.int32)
        src = torch.empty((B, T, MAXDET), device=bbox.device, dtype=torch.int32)
        _call("cvad_det_topk_decode_f32", _ptr(bbox), _ptr(conf_logit), B * T, _ptr(box), _ptr(cnt), _ptr(src), _ptr(flag), _st())
        ctx.save_for_backward(cnt, src)
        ctx.shape = bbox.shape
        ctx.mark_non_differentiable(cnt, src)
        return box, cnt, src

    @staticmethod
    def backward(ctx, dbox, _c, _s):
        cnt, src = ctx.saved_tensors
        dbox = _f32c(dbox)
        dbbox = torch.empty(ctx.shape, device=dbox.device, dtype=torch.float32)
        _call("cvad_det_topk_decode_bwd_f32", _ptr(dbox), _ptr(src), _ptr(cnt), cnt.numel(), _ptr(dbbox), _st())
        return dbbox, None, None


class _ScoreRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, pred):
        _cuda(z, pred)
        z, pred = _f32c(z), _f32c(pred)
        rows = z.numel() // 6
        out = torch.empty(z.shape[:-1] + (18,), device=z.device, dtype=torch.float32)
        _call("cvad_score_rows_f32", _ptr(z), _ptr(pred), rows, _ptr(out), _st())
        ctx.save_for_backward(z, pred)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, pred = ctx.saved_tensors
        dout = _f32c(dout)
        dz, dpred = torch.empty_like(z), torch.empty_like(pred)
        _call("cvad_score_rows_bwd_f32", _ptr(z), _ptr(pred), _ptr(dout), z.numel() // 6, _ptr(dz), _ptr(dpred), _st())
        return dz, dpred


class _MaskedMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, ntr):
        _cuda(s, ntr)
        s = _f32c(s)
        B = ntr.shape[0]
        out = torch.empty((B,), device=s.device, dtype=torch.float32)
        _call("cvad_masked_mean_f32", _ptr(s), _ptr(ntr), B, _ptr(out), _st())
        ctx.save_for_backward(ntr)
        ctx.shape = s.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        (ntr,) = ctx.saved_tensors
        dout = _f32c(dout)
        ds = torch.empty(ctx.shape, device=dout.device, dtype=torch.float32)
        _call("cvad_masked_mean_bwd_f32", _ptr(dout), _ptr(ntr), ntr.shape[0], _ptr(ds), _st())
        return ds, None


class _Ma0Loss(torch.autograd.Function):
    """vad:516-531 on (anomaly_scores (B,), kl (B,), labels (B,)) -> (total, out3 = {total, mse, kl})."""

    @staticmethod
    def forward(ctx, scores, kl, labels, flag):
        _cuda(scores, kl, labels)
        s, k = _f32c(scores), _f32c(kl)
        lab = labels.to(torch.int64).contiguous()
        out = torch.empty(3, device=s.device, dtype=torch.float32)
        ds, dk = torch.empty_like(s), torch.empty_like(k)
        _call("cvad_ma0_loss_f32", _ptr(s), _ptr(k), _ptr(lab), s.numel(), _ptr(out), _ptr(ds), _ptr(dk), _ptr(flag), _st())
        ctx.save_for_backward(ds, dk)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g, _g2):
        ds, dk = ctx.saved_tensors
        return ds * g, dk * g, None, None


def det_topk_decode(bbox, conf_logit, flag=None):
    return _DetTopkDecode.apply(bbox, conf_logit, flag)


def score_rows(z, pred):
    return _ScoreRows.apply(z, pred)


def masked_mean(s, ntr):
    return _MaskedMean.apply(s, ntr)


def ma0_loss(scores, kl, labels, flag=None):
    return _Ma0Loss.apply(scores, kl, labels, flag)


# ------------------------------------------------------------------------------------------------ modules
class PedestrianDetector(nn.Module):
    """vad:117-165.  forward returns dense (box (B,T,5,4), cnt (B,T), src (B,T,5)); at most ``num_anchors`` = 3 slots are used."""

    def __init__(self, feature_dim, num_anchors=3, max_detections=10):
        super().__init__()
        if num_anchors != 3:
            raise ValueError("the cvad_b200 M-A0 detector kernel is built for vad's 3 anchors")
        self.num_anchors, self.max_detections = num_anchors, max_detections
        self.bbox_head = nn.Linear(feature_dim, num_anchors * 4)
        self.conf_head = nn.Linear(feature_dim, num_anchors)

    def forward(self, features, flag=None):
        B, T, _ = features.shape
        bbox = ops.linear_act(features, self.bbox_head.weight, self.bbox_head.bias, ACT_NONE, None, 0.0, flag)
        # the confidences only select and order boxes (vad:148-153): nothing differentiable flows through them, conf_head.grad stays None
        with torch.no_grad():
            conf = ops.linear_act(features.detach(), self.conf_head.weight, self.conf_head.bias, ACT_NONE)
        return det_topk_decode(bbox.view(B, T, self.num_anchors, 4), conf, flag)


class AnomalyScorer(nn.Module):
    """vad:375-403: one MLP over [current | predicted | |difference|] of every track, averaged over the clip's tracks."""

    def __init__(self, num_factors):
        super().__init__()
        self.num_factors = num_factors
        self.score_net = nn.Sequential(nn.Linear(num_factors * 3, 32), nn.ReLU(), nn.Linear(32, 16), nn.ReLU(), nn.Linear(16, 1), nn.Sigmoid())

    def forward(self, z, pred, ntr):
        rows = score_rows(z, pred)                                                            # (B,5,18)
        s = _mlp(self.score_net, (0, 2, 4), rows, (ACT_RELU, ACT_RELU, ACT_SIGMOID))          # (B,5,1)
        return masked_mean(s.reshape(z.shape[0], MAXDET), ntr)


class CausalAnomalyDetector(nn.Module):
    """vad:405-454.  forward((B,T,1,H,W)) -> dict with the reference's 5 keys (+ 'dense': the batched tensors)."""

    def __init__(self, num_factors=6, reid_dim=64):
        super().__init__()
        self.backbone = ResNetBackbone(input_channels=1, output_dim=256)
        self.detector = PedestrianDetector(256 * 4 * 6)
        self.tracker = TrajectoryTracker(reid_dim=reid_dim)
        self.traj_encoder = TrajectoryEncoder(4 + reid_dim, latent_dim=32)
        self.causal_extractor = CausalFactorExtractor(32, num_factors=num_factors)
        self.structure_learner = CausalStructureLearner(num_factors)
        self.dynamics_predictor = DynamicsPredictor(num_factors)
        self.anomaly_scorer = AnomalyScorer(num_factors)
        self.noise = DeviceNoise()
        self.flags = None            # gradient-arena header (set by the trainer): activity flags for "grad is None" groups

    def set_precision(self, precision: str):
        assert precision in ("fp32", "bf16")
        self.backbone.precision = precision
        return self

    def optimizer_slots(self):
        """Parameters that may legitimately receive no gradient in a step (torch leaves .grad None, AdamW skips them): bbox_head when every
        frame fell back to the dummy box, the edge MLP when no clip has two tracks, conf_head and structure_params always."""
        slots = {}
        for p in self.detector.bbox_head.parameters():
            slots[id(p)] = SLOT_DETECTOR
        for p in self.detector.conf_head.parameters():
            slots[id(p)] = SLOT_NEVER
        for p in list(self.structure_learner.node_encoder.parameters()) + list(self.structure_learner.edge_predictor.parameters()):
            slots[id(p)] = SLOT_STRUCTURE
        slots[id(self.structure_learner.structure_params)] = SLOT_NEVER
        return slots

    def tail(self, features):
        """Everything behind the backbone: features (B,T,6144) -> the output dict (vad:426-454)."""
        B, T, _ = features.shape
        dev = features.device
        f_det = self.flags[SLOT_DETECTOR:SLOT_DETECTOR + 1] if self.flags is not None else None
        f_str = self.flags[SLOT_STRUCTURE:SLOT_STRUCTURE + 1] if self.flags is not None else None
        box, cnt, _src = self.detector(features, f_det)
        traj, ntr = self.tracker(box, cnt, f_str)
        enc = self.traj_encoder(traj, ntr)
        eps = self.noise.normal("eps", (B, MAXDET, 6), dev)     # drawn in eval mode too (vad:270-273)
        z, kl = self.causal_extractor(enc, ntr, eps)
        adj = self.structure_learner(z, ntr)
        pred = self.dynamics_predictor(z, adj)
        scores = self.anomaly_scorer(z, pred, ntr)
        dense = {"causal_factors": z, "adjacency_matrices": adj, "kl_losses": kl, "detections": box, "det_counts": cnt, "n_tracks": ntr,
                 "features": features}

        def ragged_factors():
            n = ntr.tolist()
            return [z[b, :n[b]] for b in range(B)]

        def ragged_dets():
            c = cnt.tolist()
            return [[box[b, t, :c[b][t]] for t in range(T)] for b in range(B)]

        return {
            "anomaly_scores": scores,
            "causal_factors": _LazyList(ragged_factors),
            "adjacency_matrices": _LazyList(lambda: [adj[b] for b in range(B)]),
            "kl_losses": _LazyList(lambda: [kl[b] for b in range(B)]),
            "detections": _LazyList(ragged_dets),
            "dense": dense,
        }

    def forward(self, video_frames):
        return self.tail(self.backbone(video_frames))


class MA0Trainer:
    """The loop body of vad:502-548 as an object: MSE + 0.001 KL, clip 1.0, AdamW(lr 1e-4, wd 1e-5), cosine schedule."""

    def __init__(self, model, device, num_epochs=15, lr=1e-4, precision="fp32"):
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        if self.device.type != "cuda":
            raise RuntimeError("M-A0 trainer (cvad_b200) requires a CUDA device; there is no CPU fallback")
        self.model = apply_memory_efficient_training(model).to(self.device)
        self.model.set_precision(precision)
        params = [p for p in self.model.parameters() if p.requires_grad]
        self.optimizer = FusedAdam(params, lr=lr, weight_decay=1e-5, eps=1e-8, decoupled=True, clip_mode=1, max_norm=1.0, nan_mode=1,
                                   slots=self.model.optimizer_slots())
        self.model.flags = self.optimizer.arena.header
        self.scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(self.optimizer, T_max=num_epochs)

    def loss_on_device(self, outputs, labels):
        return ma0_loss(outputs["anomaly_scores"], outputs["dense"]["kl_losses"], labels, self.optimizer.arena.header[0:1])

    def forward_backward(self, videos, labels):
        self.optimizer.zero_grad()
        outputs = self.model(videos)
        loss, comp = self.loss_on_device(outputs, labels)
        with ops.param_grad_overlap():
            loss.backward()
        return comp, outputs

    def train_step(self, videos, labels):
        comp, outputs = self.forward_backward(videos, labels)
        self.optimizer.step()
        return comp, outputs

    def mutated_tensors(self):
        a = self.optimizer.arena
        return [a.p, a.m, a.v, a.state] + [b for b in self.model.buffers()]

    def graphed_train_step(self, videos, labels):
        """One CUDA graph for zero_grad + forward + loss + backward + clip/AdamW on this batch shape."""
        from .graphs import graphed_optimizer_step

        def fwd_bwd(x, y):
            comp, out = self.forward_backward(x, y)
            return comp, out["anomaly_scores"]
        return graphed_optimizer_step(self.optimizer, fwd_bwd, (videos, labels), self.mutated_tensors())

    @torch.no_grad()
    def eval_step(self, videos, labels):
        outputs = self.model(videos)
        _, comp = self.loss_on_device(outputs, labels)
        return comp, outputs


def train_model(model, train_loader, val_loader, num_epochs=15, lr=1e-4, device="cuda", precision="bf16", verbose=True):
    """vad:474-637.  Returns (model, train_losses, val_losses).  Mixed precision = bf16 operands / fp32 accumulation on the tensor
    cores (the reference uses fp16 autocast + GradScaler, vad:491, 509); the running loss stays on the device (no per-batch .item())."""
    tr = MA0Trainer(model, device, num_epochs, lr, precision)
    train_losses, val_losses = [], []
    for epoch in range(num_epochs):
        tr.model.train()
        acc = torch.zeros(2, device=tr.device)
        for videos, labels in train_loader:
            comp, _ = tr.train_step(videos.to(tr.device, non_blocking=True), labels.to(tr.device, non_blocking=True))
            acc[0] += comp[0]
            acc[1] += 1
        tr.model.eval()
        vacc = torch.zeros(2, device=tr.device)
        for videos, labels in val_loader:
            comp, _ = tr.eval_step(videos.to(tr.device, non_blocking=True), labels.to(tr.device, non_blocking=True))
            vacc[0] += comp[0]
            vacc[1] += 1
        tr.scheduler.step()
        a, v = acc.tolist(), vacc.tolist()
        train_losses.append(a[0] / max(a[1], 1))
        val_losses.append(v[0] / max(v[1], 1))
        if verbose:
            print(f"Epoch {epoch + 1}/{num_epochs}, Train Loss: {train_losses[-1]:.6f}, Val Loss: {val_losses[-1]:.6f}")
    return tr.model, train_losses, val_losses


@torch.no_grad()
def test_model(model, test_loader, device="cuda"):
    """vad:639-657: returns (scores, labels, list of output dicts)."""
    model.eval()
    dev = torch.device(device)
    scores, labels_all, outs = [], [], []
    for videos, labels in test_loader:
        out = model(videos.to(dev))
        scores.append(out["anomaly_scores"])
        labels_all.extend(np.asarray(labels).tolist())
        outs.append(out)
    return torch.cat(scores).cpu().numpy(), np.array(labels_all), outs


# ------------------------------------------------------------------------------------------------ streaming windows
class StreamingWindowScorer:
    """Sliding-window scoring of a frame stream for the per-frame models (M-A ``ma.CausalAnomalyDetector`` and M-A0): the service form
    of the reference's batch-1 window loop (bbox:392-430).

    In eval mode the 2-D backbone treats every frame on its own (BatchNorm uses its running statistics), so a frame's 6144 features
    do not depend on the window it is scored in: they are computed ONCE, when the frame arrives, and kept in a ring on the device.  A
    window of ``clip_len`` frames every ``stride`` frames is then a gather out of the ring plus the model's tail; with stride 4 and 16-frame
    clips every frame's backbone pass -- 99 % of the forward's arithmetic -- is reused by four windows instead of being recomputed.
    ``push`` returns the scores of the windows completed by the new frames, identical to scoring each window as its own clip."""

    def __init__(self, model, clip_len=16, stride=4, capacity=256):
        if model.training:
            raise RuntimeError("StreamingWindowScorer needs model.eval(): train-mode BatchNorm couples the frames of a batch")
        if capacity < clip_len + stride:
            raise ValueError("ring capacity must hold at least one window plus one stride of new frames")
        self.model, self.clip_len, self.stride, self.capacity = model, int(clip_len), int(stride), int(capacity)
        self.ring = None            # (capacity, F) fp32 on the model's device
        self.n_frames = 0           # frames pushed so far
        self.next_window = 0        # index of the next window to score (window w starts at frame w*stride)

    @torch.no_grad()
    def push(self, frames):
        """frames (n,1,H,W) float32 (normalised) or uint8 (the loader's raw frames) on the model's device -> (window scores, first window index)."""
        if frames.dim() != 4:
            raise ValueError("expected (n, 1, H, W) frames")
        n = frames.shape[0]
        if n > self.capacity - self.clip_len:
            raise ValueError("more new frames than the ring can take while it still holds an open window")
        feats = self.model.backbone(frames.unsqueeze(0))[0]                     # (n, F): each frame once
        F = feats.shape[-1]
        if self.ring is None:
            self.ring = torch.zeros((self.capacity, F), device=feats.device, dtype=torch.float32)
        pos = self.n_frames % self.capacity
        first = min(n, self.capacity - pos)
        self.ring[pos:pos + first].copy_(feats[:first])
        if first < n:
            self.ring[:n - first].copy_(feats[first:])
        self.n_frames += n
        n_win = 0
        if self.n_frames >= self.clip_len:
            n_win = (self.n_frames - self.clip_len) // self.stride + 1 - self.next_window
        w0 = self.next_window
        if n_win <= 0:
            return torch.empty(0, device=feats.device), w0
        clips = torch.empty((n_win, self.clip_len, F), device=feats.device, dtype=torch.float32)
        _call("cvad_window_features_f32", _ptr(self.ring), self.capacity, w0 * self.stride, self.stride, self.clip_len, F, n_win, _ptr(clips),
              _st())
        self.next_window += n_win
        return self.model.tail(clips)["anomaly_scores"], w0

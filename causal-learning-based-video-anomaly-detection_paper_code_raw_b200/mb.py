"""M-B: the checkpointed compact 3-D CNN + NOTEARS-style causal head and its trainers, on the cvad_b200 kernels.

Mirrors avenue_training_script2.py: ``CompactFeatureExtractor`` s2:15-35, ``DifferentiableCausalDiscovery`` s2:37-67,
``CausalAnomalyDetector`` s2:69-101, ``ImprovedMiniCausalVAD`` s2:107-297, plus the ``MiniCausalVAD`` facade that
avenue_training_script1.py imports from the (missing) ``minicausal_vad`` module (call sites s1:43-51, 101-106, 141,
151-154, 161, 184-188, 199-205, 299-306).  Constructor signatures, attribute names, ``state_dict`` keys / shapes
(so best_improved_model.pth loads with strict=True) and return conventions are the reference's; the arithmetic is not
torch's: every layer calls the C ABI in include/cvad_b200.h.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .arena import FusedAdam
from .noise import DeviceNoise
from .ops import ACT_NONE, ACT_RELU, ACT_SIGMOID


class CompactFeatureExtractor(nn.Module):
    """s2:15-35.  The nn.Conv3d / nn.Linear members only hold parameters (names, shapes, default init)."""

    def __init__(self, input_channels=3, feature_dim=64):
        super().__init__()
        self.conv3d_1 = nn.Conv3d(input_channels, 16, (3, 3, 3), stride=(1, 2, 2), padding=1)
        self.conv3d_2 = nn.Conv3d(16, 32, (3, 3, 3), stride=(2, 2, 2), padding=1)
        self.conv3d_3 = nn.Conv3d(32, 64, (3, 3, 3), stride=(2, 2, 2), padding=1)
        self.adaptive_pool = nn.AdaptiveAvgPool3d((4, 4, 4))
        self.fc = nn.Linear(64 * 4 * 4 * 4, feature_dim)
        self.dropout = nn.Dropout(0.3)
        self.noise = DeviceNoise()

    def forward(self, x):
        x = ops.conv_act(x, self.conv3d_1.weight, self.conv3d_1.bias, (1, 2, 2), 1, ACT_RELU)
        x = ops.conv_act(x, self.conv3d_2.weight, self.conv3d_2.bias, 2, 1, ACT_RELU)
        x = ops.conv_act(x, self.conv3d_3.weight, self.conv3d_3.bias, 2, 1, ACT_RELU)
        x = ops.adaptive_avgpool(x, (4, 4, 4))
        x = x.reshape(x.size(0), -1)
        keep = None
        if self.training:   # Dropout(0.3) on the feature itself (s2:34), fused into the fc epilogue
            keep = self.noise.keep_mask("feat", (x.size(0), self.fc.out_features), self.dropout.p, x.device)
        return ops.linear_act(x, self.fc.weight, self.fc.bias, ACT_NONE, keep, self.dropout.p)


class DifferentiableCausalDiscovery(nn.Module):
    """s2:37-67: Linear(16,32)+ReLU+Linear(32,256)+Sigmoid, zero diagonal."""

    def __init__(self, num_variables=16, hidden_dim=32):
        super().__init__()
        self.num_variables = num_variables
        self.causal_net = nn.Sequential(nn.Linear(num_variables, hidden_dim), nn.ReLU(),
                                        nn.Linear(hidden_dim, num_variables * num_variables), nn.Sigmoid())
        self._offdiag = {}

    def _mask(self, B, device):
        key = (B, str(device))
        m = self._offdiag.get(key)
        if m is None:
            v = self.num_variables
            m = (1.0 - torch.eye(v, device=device)).reshape(1, v * v).expand(B, v * v).contiguous()
            self._offdiag = {key: m}
        return m

    def forward(self, features):
        B = features.size(0)
        h = ops.linear_act(features, self.causal_net[0].weight, self.causal_net[0].bias, ACT_RELU)
        # sigmoid and the (1 - I) self-loop mask (s2:57-58) share the GEMM epilogue
        a = ops.linear_act(h, self.causal_net[2].weight, self.causal_net[2].bias, ACT_SIGMOID, self._mask(B, features.device), 0.0)
        return a.view(B, self.num_variables, self.num_variables)

    def acyclicity_constraint(self, adj_matrix):
        """s2:62-67 (unused by the improved loss, kept for API parity): trace((mean_b A + 1e-8)^2)."""
        abar = adj_matrix.mean(dim=0) + 1e-8
        return torch.trace(abar @ abar)


class CausalAnomalyDetector(nn.Module):
    """s2:69-101.  forward(video_clips (B,3,T,H,W)) -> (anomaly_scores (B,1), causal_adj (B,16,16), features (B,16))."""

    def __init__(self, feature_dim=64, causal_dim=16, hidden_dim=128):
        super().__init__()
        self.feature_extractor = CompactFeatureExtractor(feature_dim=causal_dim)
        self.causal_discovery = DifferentiableCausalDiscovery(num_variables=causal_dim)
        self.graph_encoder = nn.Sequential(nn.Linear(causal_dim * causal_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.3),
                                           nn.Linear(hidden_dim, 64))
        self.anomaly_predictor = nn.Sequential(nn.Linear(causal_dim + 64, 32), nn.ReLU(), nn.Linear(32, 1), nn.Sigmoid())

    @property
    def noise(self):
        return self.feature_extractor.noise

    @noise.setter
    def noise(self, n):
        self.feature_extractor.noise = n

    def forward(self, video_clips):
        features = self.feature_extractor(video_clips)
        causal_adj = self.causal_discovery(features)
        B = causal_adj.size(0)
        ge, ap = self.graph_encoder, self.anomaly_predictor
        keep = None
        if self.training:
            keep = self.noise.keep_mask("graph", (B, ge[0].out_features), ge[2].p, features.device)
        g = ops.linear_act(causal_adj.view(B, -1), ge[0].weight, ge[0].bias, ACT_RELU, keep, ge[2].p)
        g = ops.linear_act(g, ge[3].weight, ge[3].bias)
        combined = torch.cat([features, g], dim=1)
        s = ops.linear_act(combined, ap[0].weight, ap[0].bias, ACT_RELU)
        s = ops.linear_act(s, ap[2].weight, ap[2].bias, ACT_SIGMOID)
        return s, causal_adj, features


COMPONENT_KEYS = ("anomaly_loss", "acyclicity_loss", "sparsity_loss", "consistency_loss", "structure_loss", "edge_count",
                  "sparsity_ratio")


class ImprovedMiniCausalVAD:
    """s2:107-297 with the same attributes and method signatures.  Differences are only in *where* things run: the
    5-term loss is one fused kernel pair, clip+AdamW is one fused arena step, the NaN-skip is a device flag and the
    loss components are accumulated on the device and read back once per epoch (the reference syncs 9x per batch)."""

    def __init__(self, device="cuda", lr=0.0005, weight_decay=0.001, verbose=True, dp=None):
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        if self.device.type != "cuda":
            raise RuntimeError("ImprovedMiniCausalVAD (cvad_b200) requires a CUDA device; there is no CPU fallback")
        self.model = CausalAnomalyDetector().to(self.device)
        self.optimizer = FusedAdam(self.model.parameters(), lr=lr, weight_decay=weight_decay, decoupled=True,
                                   clip_mode=1, max_norm=0.5, nan_mode=1)
        self.anomaly_weight = 1.0
        self.causal_weight = 0.01
        self.sparsity_weight = 0.001
        self.consistency_weight = 0.01
        self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, mode="min", factor=0.5, patience=5)
        self.dp = dp
        if dp is not None:
            dp.attach(self.optimizer)
        self.verbose = verbose
        if verbose:
            print(f"ImprovedMiniCausalVAD initialized on {self.device}")
            print(f"Model parameters: {sum(p.numel() for p in self.model.parameters()):,}")

    # -- loss --------------------------------------------------------------------------------------------------
    def _weights(self):
        return (self.anomaly_weight, self.causal_weight, self.sparsity_weight, self.consistency_weight)

    def loss_on_device(self, anomaly_scores, causal_adj, targets, pseudo_targets=None):
        """Returns (loss 0-d tensor, components (8,) device tensor = [total] + COMPONENT_KEYS).  No host sync."""
        if pseudo_targets is None:   # s2:139-141: 5 % random anomalies, the true labels are ignored
            u = self.model.noise.uniform("pseudo", tuple(targets.shape), targets.device)
            pseudo_targets = (u > 0.95).float()
        flag = self.optimizer.arena.header[0:1]
        return ops.mb_loss(anomaly_scores, causal_adj, pseudo_targets, self._weights(), flag)

    def compute_improved_loss(self, anomaly_scores, causal_adj, targets, features):
        loss, comp = self.loss_on_device(anomaly_scores, causal_adj, targets)
        vals = comp.tolist()   # API parity with s2:197-205 (python floats) -> this call synchronises
        return loss, dict(zip(COMPONENT_KEYS, vals[1:]))

    # -- training ----------------------------------------------------------------------------------------------
    def train_step(self, videos, labels, pseudo_targets=None):
        """One iteration of s2:217-238 on tensors already on the device.  Returns the (8,) component tensor."""
        self.optimizer.zero_grad()
        scores, adj, _ = self.model(videos)
        loss, comp = self.loss_on_device(scores, adj, labels, pseudo_targets)
        loss.backward()
        self.optimizer.step()
        return comp

    def graphed_train_step(self, videos, labels):
        """The same iteration as one CUDA graph (two around the gradient all-reduce under data parallelism): ``(videos, labels) ->
        ((8,) component tensor,)``.  Pseudo-labels are drawn on the device inside the graph (s2:139-141)."""
        from .graphs import graphed_optimizer_step

        def fwd_bwd(x, y):
            self.optimizer.zero_grad()
            scores, adj, _ = self.model(x)
            loss, comp = self.loss_on_device(scores, adj, y)
            loss.backward()
            return (comp,)
        a = self.optimizer.arena
        return graphed_optimizer_step(self.optimizer, fwd_bwd, (videos, labels), [a.p, a.m, a.v, a.state])

    def train_epoch_improved(self, dataloader):
        self.model.train()
        acc = torch.zeros(8, device=self.device)
        for batch_idx, (videos, labels) in enumerate(dataloader):
            videos = videos.to(self.device, non_blocking=True)
            labels = labels.to(self.device, non_blocking=True).float()
            comp = self.train_step(videos, labels)
            # a NaN loss is skipped by the optimizer kernel (device flag); keep it out of the running sums too (s2:230-232)
            acc += torch.nan_to_num(comp, nan=0.0, posinf=0.0, neginf=0.0)
        num_batches = max(len(dataloader), 1)
        vals = (acc / num_batches).tolist()          # the single host sync of the epoch
        avg_loss = vals[0]
        avg_components = dict(zip(COMPONENT_KEYS, vals[1:]))
        self.scheduler.step(avg_loss)
        return avg_loss, avg_components

    @torch.no_grad()
    def evaluate_improved(self, dataloader, return_arrays: bool = True):
        """s2:265-297.  The eight metrics are computed on the device (evaltail.mb_eval_metrics: one read-back of 8 scalars); the score and
        graph arrays the reference returns are copied to the host only when ``return_arrays`` (the reference's callers s2:426 keep the
        predictions; pass False to leave the (N,16,16) graphs on the device and get ``(None, None, metrics)``)."""
        from . import evaltail
        self.model.eval()
        preds, graphs = [], []
        for videos, _ in dataloader:
            videos = videos.to(self.device, non_blocking=True)
            s, a, _f = self.model(videos)
            preds.append(s.reshape(-1))
            graphs.append(a)
        p_dev, g_dev = torch.cat(preds), torch.cat(graphs)
        vals = evaltail.mb_eval_metrics(p_dev, g_dev, 0.1).tolist()
        eval_metrics = dict(zip(evaltail.METRIC_KEYS, vals))
        eval_metrics["unique_graphs"] = int(eval_metrics["unique_graphs"])
        if not return_arrays:
            return None, None, eval_metrics
        return p_dev.cpu().numpy(), g_dev.cpu().numpy(), eval_metrics


class MiniCausalVAD(ImprovedMiniCausalVAD):
    """The object avenue_training_script1.py drives (s1:101): same model, lr 1e-3 default (s1:104), 4-key components."""

    def __init__(self, device="cuda", verbose=False, dp=None):
        super().__init__(device=device, lr=0.001, verbose=verbose, dp=dp)

    def train_epoch(self, dataloader):
        loss, comps = self.train_epoch_improved(dataloader)
        keep = ("anomaly_loss", "acyclicity_loss", "sparsity_loss", "consistency_loss")
        return loss, {k: comps[k] for k in keep}

    def evaluate(self, dataloader):
        predictions, graphs, metrics = self.evaluate_improved(dataloader)
        return predictions, metrics, graphs

    def save_model(self, path: str):
        torch.save({"model_state_dict": self.model.state_dict(), "optimizer_state_dict": self.optimizer.state_dict()}, path)

    def load_model(self, path: str):
        ck = torch.load(path, map_location=self.device, weights_only=False)
        sd = ck.get("model_state_dict", ck.get("state_dict", ck)) if isinstance(ck, dict) else ck
        self.model.load_state_dict(sd, strict=True)
        if isinstance(ck, dict) and "optimizer_state_dict" in ck:
            self.optimizer.load_state_dict(ck["optimizer_state_dict"])


@torch.no_grad()
def create_unsupervised_labels(test_loader, model, threshold_percentile=95):
    """avenue_training_script1.py:36-67: scores of the whole loader, their percentile threshold and the pseudo-labels above it.
    ``model`` is a MiniCausalVAD-style trainer (``.model`` callable returning a 3-tuple, ``.device``).  Scores stay on the
    device until one concatenated read-back (the reference copies every batch)."""
    from . import evaltail
    model.model.eval()
    parts = []
    for videos, _ in test_loader:
        s, _, _ = model.model(videos.to(model.device, non_blocking=True))
        parts.append(s.reshape(-1))
    if not parts:
        return np.zeros(0, dtype=np.float32), np.zeros(0), 0.0
    scores = torch.cat(parts)
    thr, labels = evaltail.percentile_labels(scores, float(threshold_percentile))      # np.percentile + comparison, on the device
    return scores.cpu().numpy(), labels.cpu().numpy().astype(float), float(thr)

"""M-C: ``SimpleVideoAnomalyDetector`` and ``StableTrainer`` on the cvad_b200 kernels.

Mirrors minicausal_vad_complete3.py: model mc3:25-102 (3x [Conv3d k3 p1 + BatchNorm3d + ReLU + MaxPool3d], global average
pool, Dropout/Linear classifier with sigmoid), trainer mc3:218-431 (BCE, Adam with L2 decay, clip to 1.0 only when the
gradient norm exceeds 10, StepLR(15, 0.7), best-AUC checkpoint ``{'model_state_dict','epoch','best_auc'}``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .arena import FusedAdam
from .noise import DeviceNoise
from .ops import ACT_NONE, ACT_RELU, ACT_SIGMOID

_POOLS = ((1, 2, 2), (2, 2, 2), (2, 2, 2))


class SimpleVideoAnomalyDetector(nn.Module):
    def __init__(self, input_channels=1, temporal_frames=8, spatial_size=64):
        super().__init__()
        self.temporal_frames = temporal_frames
        self.spatial_size = spatial_size
        layers = []
        cin = input_channels
        for cout, pool in zip((8, 16, 32), _POOLS):
            layers += [nn.Conv3d(cin, cout, kernel_size=3, stride=1, padding=1), nn.BatchNorm3d(cout), nn.ReLU(inplace=True),
                       nn.MaxPool3d(kernel_size=pool, stride=pool)]
            cin = cout
        layers.append(nn.AdaptiveAvgPool3d((1, 1, 1)))
        self.features = nn.Sequential(*layers)                      # indices 0,1,4,5,8,9 hold parameters (mc3:36-57)
        self.classifier = nn.Sequential(nn.Dropout(0.5), nn.Linear(32, 16), nn.ReLU(inplace=True), nn.Dropout(0.3),
                                        nn.Linear(16, 8), nn.ReLU(inplace=True), nn.Linear(8, 1), nn.Sigmoid())
        self._initialize_weights()
        self.to(dtype=torch.float32)                                # mc3:74
        self.noise = DeviceNoise()

    def _initialize_weights(self):
        """mc3:76-88: He-normal (fan_out) convs, unit BN, N(0, 0.01) linears, zero biases."""
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        if len(x.shape) != 5:
            raise ValueError(f"Expected 5D tensor (B,C,T,H,W), got {x.shape}")     # mc3:92-93
        h = x
        for blk, pool in enumerate(_POOLS):
            conv, bn = self.features[4 * blk], self.features[4 * blk + 1]
            if not (self.training and bn.training) and not torch.is_grad_enabled():
                # inference: the BatchNorm's running statistics fold into the convolution (one launch for conv + BN + ReLU)
                wf, bf = ops.bn_fold_conv(conv.weight, conv.bias, bn)
                h = ops.maxpool(ops.conv_act(h, wf, bf, 1, 1, ACT_RELU), pool, pool, 0)
                continue
            h = ops.conv_act(h, conv.weight, conv.bias, 1, 1, ACT_NONE)
            h = ops.batchnorm_act(h, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                  ops.bn_workspace(h.device, bn.num_features), self.training and bn.training, ACT_RELU, bn.eps,
                                  bn.momentum)
            h = ops.maxpool(h, pool, pool, 0)
        h = ops.adaptive_avgpool(h, (1, 1, 1))
        f = h.reshape(h.size(0), -1)
        cl = self.classifier
        B = f.size(0)
        keep1 = None
        if self.training:
            f = ops.mask_scale(f, self.noise.keep_mask("cls0", (B, 32), cl[0].p, f.device), cl[0].p)
            keep1 = self.noise.keep_mask("cls1", (B, 16), cl[3].p, f.device)
        h = ops.linear_act(f, cl[1].weight, cl[1].bias, ACT_RELU, keep1, cl[3].p)
        h = ops.linear_act(h, cl[4].weight, cl[4].bias, ACT_RELU)
        return ops.linear_act(h, cl[6].weight, cl[6].bias, ACT_SIGMOID)


def roc_auc(targets: np.ndarray, scores: np.ndarray) -> float:
    """Rank-based ROC-AUC with tie handling (what sklearn.metrics.roc_auc_score returns, mc3:388)."""
    t = np.asarray(targets).astype(bool)
    s = np.asarray(scores, dtype=np.float64)
    npos, nneg = int(t.sum()), int((~t).sum())
    if npos == 0 or nneg == 0:
        return 0.0
    order = np.argsort(s, kind="mergesort")
    ranks = np.empty(len(s), dtype=np.float64)
    ss = s[order]
    i = 0
    while i < len(ss):
        j = i
        while j + 1 < len(ss) and ss[j + 1] == ss[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return float((ranks[t].sum() - npos * (npos + 1) / 2.0) / (npos * nneg))


class StableTrainer:
    """mc3:218-431.  ``train_epoch() -> (avg_loss, accuracy)``, ``evaluate() -> (avg_loss, auc, accuracy)``."""

    def __init__(self, model, train_loader, test_loader, device, lr=0.001, dp=None):
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.model = model.to(self.device)
        self.train_loader, self.test_loader = train_loader, test_loader
        self.optimizer = FusedAdam(self.model.parameters(), lr=lr, weight_decay=1e-5, eps=1e-8, decoupled=False, clip_mode=2,
                                   max_norm=1.0, clip_threshold=10.0, nan_mode=1)
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=15, gamma=0.7)
        self.history = {"train_loss": [], "test_loss": [], "test_auc": [], "train_acc": [], "test_acc": []}
        self.best_auc = 0.0
        if dp is not None:
            dp.attach(self.optimizer)

    def train_step(self, data, targets):
        """mc3:269-311 on device tensors; returns (loss 0-d tensor, scores (B,))."""
        self.optimizer.zero_grad()
        out = self.model(data).reshape(-1)
        loss = ops.bce_loss(out, targets, self.optimizer.arena.header[0:1])
        loss.backward()
        self.optimizer.step()
        return loss.detach(), out.detach()

    def train_epoch(self):
        self.model.train()
        acc = torch.zeros(3, device=self.device)        # loss sum, correct, total
        for data, targets in self.train_loader:
            data = data.to(dtype=torch.float32).to(self.device, non_blocking=True)
            targets = targets.to(dtype=torch.float32).to(self.device, non_blocking=True)
            loss, out = self.train_step(data, targets)
            ok = torch.isfinite(loss).float()
            acc[0] += torch.nan_to_num(loss) * ok
            acc[1] += ((out > 0.5).float() == targets).float().sum() * ok
            acc[2] += targets.numel() * ok
        vals = acc.tolist()
        n = len(self.train_loader)
        return (vals[0] / n if n > 0 else 0), (vals[1] / vals[2] if vals[2] > 0 else 0)

    @torch.no_grad()
    def evaluate(self):
        self.model.eval()
        outs, tgts, losses = [], [], []
        for data, targets in self.test_loader:
            data = data.float().to(self.device, non_blocking=True)
            targets = targets.float().to(self.device, non_blocking=True)
            out = self.model(data).reshape(-1)
            losses.append(ops.bce_loss(out, targets))
            outs.append(out)
            tgts.append(targets)
        if not outs:
            return float("inf"), 0.0, 0
        # mc3:372-390 on the device: non-finite outputs are dropped, accuracy at 0.5, rank-based ROC-AUC (evaltail.roc_auc);
        # one read-back of four scalars instead of every output and target
        from . import evaltail
        o, t, ls = torch.cat(outs), torch.cat(tgts), torch.stack(losses)
        good = torch.isfinite(o)
        og, tg = o[good], t[good]
        finite_ls = torch.where(torch.isfinite(ls), ls, torch.zeros_like(ls)).sum().double().reshape(1)
        correct = ((og > 0.5) == (tg > 0.5)).double().sum().reshape(1)
        auc = evaltail.roc_auc(og, tg) if og.numel() else torch.zeros(1, device=o.device, dtype=torch.float64)
        vals = torch.cat([finite_ls, correct, auc, torch.tensor([float(og.numel())], device=o.device, dtype=torch.float64)]).tolist()
        avg_loss = vals[0] / len(self.test_loader)
        acc = vals[1] / vals[3] if vals[3] > 0 else 0
        return avg_loss, (vals[2] if vals[3] > 0 else 0.0), acc

    def train_model(self, epochs, save_path="simple_anomaly_model.pth"):
        for epoch in range(epochs):
            train_loss, train_acc = self.train_epoch()
            test_loss, test_auc, test_acc = self.evaluate()
            self.scheduler.step()
            for k, v in zip(("train_loss", "test_loss", "test_auc", "train_acc", "test_acc"),
                            (train_loss, test_loss, test_auc, train_acc, test_acc)):
                self.history[k].append(v)
            if test_auc > self.best_auc:
                self.best_auc = test_auc
                torch.save({"model_state_dict": self.model.state_dict(), "epoch": epoch, "best_auc": self.best_auc}, save_path)
            if epoch > 20 and test_auc < 0.55 and train_loss < 0.1:     # mc3:427-429
                break

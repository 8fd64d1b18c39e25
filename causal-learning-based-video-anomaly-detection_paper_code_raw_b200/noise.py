"""Noise sources for the stochastic sites of the hot path (dropout keep-masks, pseudo-label draws, VAE eps).

Production draws on the device; parity tests inject the exact tensors the oracle / reference consumed
(SURVEY.md section 7 "RNG parity": s2:141 rand_like, cad:330 randn_like in eval too, 2+6+2 Dropout sites).
"""
from __future__ import annotations

import torch


class DeviceNoise:
    """Default: draw from torch's CUDA generator (seedable with torch.manual_seed / a per-rank seed)."""

    def __init__(self, generator: torch.Generator | None = None):
        self.generator = generator

    def keep_mask(self, name: str, shape, p: float, device) -> torch.Tensor:
        # one launch: Bernoulli(1 - p) keep flags as fp32 (torch.rand(...) >= p followed by .float() is three)
        return torch.empty(shape, device=device, dtype=torch.float32).bernoulli_(1.0 - p, generator=self.generator)

    def uniform(self, name: str, shape, device) -> torch.Tensor:
        return torch.rand(shape, device=device, generator=self.generator)

    def normal(self, name: str, shape, device) -> torch.Tensor:
        return torch.randn(shape, device=device, generator=self.generator)


class FixedNoise(DeviceNoise):
    """Replays caller-supplied tensors: ``values[name]`` is a tensor or a FIFO list of tensors."""

    def __init__(self, values: dict):
        super().__init__()
        self.values = {k: (list(v) if isinstance(v, (list, tuple)) else [v]) for k, v in values.items()}

    def _pop(self, name, shape, device):
        q = self.values.get(name)
        if not q:
            raise KeyError(f"FixedNoise has no tensor queued for '{name}'")
        t = q.pop(0) if len(q) > 1 else q[0]
        t = t.to(device=device, dtype=torch.float32)
        if tuple(t.shape) != tuple(shape):
            t = t.reshape(shape)
        return t

    def keep_mask(self, name, shape, p, device):
        return self._pop(name, shape, device)

    def uniform(self, name, shape, device):
        return self._pop(name, shape, device)

    def normal(self, name, shape, device):
        return self._pop(name, shape, device)

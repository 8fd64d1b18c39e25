"""Device-side frame path: a video's raw uint8 frames live on the GPU; resize, colour order, scaling and clip windowing
happen there.

The reference's loaders do this per clip on the host (``cv2.imread`` -> ``cv2.resize`` -> float -> normalise -> stack: cad:89-104,
the Avenue loaders behind s1:19 / s2:357-365, the sliding windows of bbox:392-411) and ship fp32 clips over PCIe; consecutive clips share
all but ``stride`` of their frames, so every frame is decoded, resized and copied 2-4 times.  Here a video is uploaded ONCE as bytes
(``DeviceFrames``, pinned staging + one async copy), resized once with a kernel that reproduces ``cv2.resize`` (INTER_LINEAR, 8-bit) bit
for bit, and clips are gathered by start index:

    frames = DeviceFrames(video_u8, device).resized((64, 64))
    clips = frames.clips_f32(starts, T=8)             # (B,3,T,64,64) float in [0,1], RGB: what create_avenue_dataloaders yields
    clips = frames.clips_u8(starts, T=16)             # (B,T,1,H,W) uint8 for M-A (normalised inside the stem)
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .ops import _call, _ptr, _st


class DeviceFrames:
    def __init__(self, frames, device="cuda"):
        """``frames``: (F,H,W) grayscale or (F,H,W,C) interleaved uint8, numpy or torch (host or device)."""
        t = torch.from_numpy(np.ascontiguousarray(frames)) if isinstance(frames, np.ndarray) else frames
        if t.dtype != torch.uint8:
            raise TypeError("DeviceFrames holds raw uint8 frames (what cv2.imread returns)")
        if t.dim() == 3:
            t = t.unsqueeze(-1)
        if t.dim() != 4 or not 1 <= t.shape[-1] <= 4:
            raise ValueError(f"expected (F,H,W) or (F,H,W,C<=4) frames, got {tuple(t.shape)}")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DeviceFrames needs a CUDA device (there is no CPU fallback)")
        if not t.is_cuda:
            t = t.contiguous()
            t = (t if t.is_pinned() else t.pin_memory()).to(dev, non_blocking=True)      # one upload per video
        self.data = t.contiguous()

    @property
    def shape(self):
        return tuple(self.data.shape)

    def __len__(self):
        return self.data.shape[0]

    def resized(self, size) -> "DeviceFrames":
        """``cv2.resize(frame, size)`` (size = (width, height), INTER_LINEAR) of every frame, bit-exact."""
        F, H, W, C = self.data.shape
        dw, dh = int(size[0]), int(size[1])
        if (dw, dh) == (W, H):
            return self
        out = torch.empty((F, dh, dw, C), device=self.data.device, dtype=torch.uint8)
        _call("cvad_resize_bilinear_u8", _ptr(self.data), F, H, W, C, _ptr(out), dh, dw, _st())
        return DeviceFrames(out, self.data.device)

    def _starts(self, starts):
        s = torch.as_tensor(starts, dtype=torch.int32)
        return s.to(self.data.device, non_blocking=True).contiguous()

    def clips_f32(self, starts, T: int, frame_stride: int = 1, scale: float = 1.0 / 255.0, rgb: bool = True) -> torch.Tensor:
        """(B,C,T,H,W) fp32 clips starting at ``starts``: value * scale, BGR -> RGB when ``rgb`` (frames as cv2.imread delivers them)."""
        F, H, W, C = self.data.shape
        s = self._starts(starts)
        out = torch.empty((s.numel(), C, T, H, W), device=self.data.device, dtype=torch.float32)
        _call("cvad_clips_from_frames_f32", _ptr(self.data), F, H, W, C, _ptr(s), s.numel(), T, frame_stride, float(scale), int(bool(rgb) and C == 3),
              _ptr(out), _st())
        return out

    def clips_u8(self, starts, T: int, frame_stride: int = 1) -> torch.Tensor:
        """(B,T,1,H,W) uint8 clips of a grayscale video: M-A's input (cad:57, 89-96), normalised on the fly by the stem."""
        F, H, W, C = self.data.shape
        if C != 1:
            raise ValueError("clips_u8 expects grayscale frames")
        s = self._starts(starts)
        out = torch.empty((s.numel(), T, 1, H, W), device=self.data.device, dtype=torch.uint8)
        _call("cvad_clips_from_frames_u8", _ptr(self.data), F, H, W, _ptr(s), s.numel(), T, frame_stride, _ptr(out), _st())
        return out


def window_starts(n_frames: int, window: int = 8, stride: int = 4):
    """bbox:392: ``range(0, n_frames - window, stride)``."""
    return list(range(0, n_frames - window, stride))

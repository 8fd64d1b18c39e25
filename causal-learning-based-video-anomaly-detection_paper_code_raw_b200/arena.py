"""Flat fp32 parameter / gradient / Adam-moment arenas and the fused clip+Adam(W) step (K12, K13 of SURVEY.md 2.5).

Memory layout in HBM (all fp32, one allocation each for p, g, m, v):

    [ header: 1024 floats | tensor 0 padded to 1024 | tensor 1 padded to 1024 | ... ]

* ``param.data`` and ``param.grad`` of every managed parameter are views into the p / g arenas, so weight-gradient
  kernels write straight into the buffer that the NCCL all-reduce and the optimizer kernel read -- no pack/unpack.
* g-header[0] is the device-side "loss was NaN/Inf" flag, g-header[1..7] are activity flags of parameter groups that
  may legitimately receive no gradient in a step (torch leaves ``grad=None`` and AdamW skips them: SURVEY fact 6).
* One step = ``cvad_sumsq_f32`` (grad-norm^2 + non-finite test) + ``cvad_adam_flat_f32`` (clip, moment update, decay,
  skip logic) -- no ``.item()``, no host round trip (the reference syncs at s2:230, s2:241, mc3:298-305).

``FusedAdam`` is a ``torch.optim.Optimizer`` so schedulers (ReduceLROnPlateau s2:128, CosineAnnealingLR cad:620, StepLR
mc3:237) and ``state_dict()`` keep working, and its state_dict has the stock AdamW layout of best_improved_model.pth.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from ._lib import OptState

BLOCK = 1024
HEADER = 1024


class FlatArena:
    def __init__(self, params, slots=None):
        params = [p for p in params]
        assert params, "no parameters"
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatArena needs CUDA parameters: move the model to the GPU first (no CPU fallback)")
        self.params = params
        self.device = dev
        slots = slots or {}
        self.offsets = []
        off = HEADER
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + BLOCK - 1) // BLOCK * BLOCK
        self.total = off
        self.p = torch.zeros(off, device=dev, dtype=torch.float32)
        self.g = torch.zeros(off, device=dev, dtype=torch.float32)
        self.m = torch.zeros(off, device=dev, dtype=torch.float32)
        self.v = torch.zeros(off, device=dev, dtype=torch.float32)
        bs = torch.full((off // BLOCK,), -1, dtype=torch.int32)
        self.slot_of = []
        for p, o in zip(params, self.offsets):
            n = p.numel()
            slot = int(slots.get(id(p), 0))
            self.slot_of.append(slot)
            bs[o // BLOCK: (o + n + BLOCK - 1) // BLOCK] = slot
            view = self.p[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.g[o:o + n].view(p.shape)
        self.block_slot = bs.to(dev)
        self.state = torch.zeros(ctypes.sizeof(OptState), device=dev, dtype=torch.uint8)
        self.header = self.g[:16]

    # -- views ------------------------------------------------------------------------------------------
    def view(self, buf, i):
        p, o = self.params[i], self.offsets[i]
        return buf[o:o + p.numel()].view(p.shape)

    def zero_grad(self):
        from .ops import _call, _ptr, _st
        _call("cvad_fill_f32", _ptr(self.g), self.g.numel(), 0.0, _st())
        for p, o in zip(self.params, self.offsets):   # re-attach if something detached the views (e.g. set_to_none)
            if p.grad is None or p.grad.data_ptr() != self.g.data_ptr() + 4 * o:
                p.grad = self.g[o:o + p.numel()].view(p.shape)

    def mark_active(self, slot: int):
        """Host-side activation of a group flag (device-side kernels may also set it)."""
        ops._call("cvad_fill_f32", self.g.data_ptr() + 4 * slot, 1, 1.0, ops._st())

    def read_state(self) -> OptState:
        raw = bytes(self.state.cpu().numpy().tobytes())
        return OptState.from_buffer_copy(raw)

    def write_steps(self, steps):
        st = self.read_state()
        for k in range(8):
            st.step[k] = int(steps[k])
        buf = torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8)
        self.state.copy_(buf.to(self.device))

    def set_device_lr(self, lr: float):
        """Keep the learning rate in device memory so a captured CUDA graph follows the scheduler (0 = use the argument)."""
        if getattr(self, "_device_lr", None) == lr:
            return
        off = OptState.lr_device.offset
        self.state[off:off + 8].copy_(torch.tensor([lr], dtype=torch.float64).view(torch.uint8), non_blocking=False)
        self._device_lr = lr

    def step(self, lr, betas, eps, weight_decay, decoupled, clip_mode, max_norm, clip_threshold, nan_mode, grad_scale=1.0):
        st = ops._st()
        ops._call("cvad_sumsq_f32", self.g.data_ptr() + 4 * HEADER, self.total - HEADER, float(grad_scale), self.state.data_ptr(), st)
        ops._call("cvad_adam_flat_f32", self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.total,
                  self.block_slot.data_ptr(), self.state.data_ptr(), float(grad_scale), float(lr), float(betas[0]), float(betas[1]),
                  float(eps), float(weight_decay), int(decoupled), int(clip_mode), float(max_norm), float(clip_threshold),
                  int(nan_mode), st)


class FusedAdam(torch.optim.Optimizer):
    """Adam / AdamW over a FlatArena with fused global-norm clipping and on-device skip logic.

    clip_mode: 0 none | 1 clip to ``max_norm`` always (s2:236, cad:666) | 2 clip to ``max_norm`` only when the norm
    exceeds ``clip_threshold`` (mc3:308-309).  nan_mode 1 skips the whole step on a non-finite loss/gradient
    (s2:230-232; GradScaler semantics cad:667; mc3:301-304 under set_to_none zero_grad).
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, decoupled=True, clip_mode=0,
                 max_norm=1.0, clip_threshold=0.0, nan_mode=1, slots=None):
        params = [p for p in params]
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, foreach=None, maximize=False,
                        capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self.arena = FlatArena(params, slots)
        self.decoupled, self.clip_mode, self.max_norm = decoupled, clip_mode, max_norm
        self.clip_threshold, self.nan_mode = clip_threshold, nan_mode
        self.grad_scale = 1.0
        self.pre_step_hook = None          # e.g. the data-parallel gradient all-reduce
        for i, p in enumerate(params):
            self.state[p] = {"step": torch.tensor(0.0), "exp_avg": self.arena.view(self.arena.m, i),
                             "exp_avg_sq": self.arena.view(self.arena.v, i)}

    def zero_grad(self, set_to_none: bool = True):   # noqa: ARG002 - grads live in the arena, they are zeroed not dropped
        self.arena.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        if self.pre_step_hook is not None:
            self.pre_step_hook(self.arena)
        return self.step_local()

    @torch.no_grad()
    def step_local(self):
        """The fused grad-norm + clip + Adam(W) kernels alone (no data-parallel hook): what runs after the all-reduce."""
        g = self.param_groups[0]
        if getattr(self.arena, "_device_lr", None) is not None and not torch.cuda.is_current_stream_capturing():
            # a captured step published the learning rate to the device (lr_device overrides the argument from then on): keep it in
            # step with the scheduler for eager steps as well (cached: a copy only when the value changed)
            self.arena.set_device_lr(float(g["lr"]))
        self.arena.step(g["lr"], g["betas"], g["eps"], g["weight_decay"], self.decoupled, self.clip_mode, self.max_norm,
                        self.clip_threshold, self.nan_mode, self.grad_scale)
        return None

    # -- stock-format (de)serialisation ---------------------------------------------------------------------
    def _sync_steps(self):
        st = self.arena.read_state()
        for p, slot in zip(self.arena.params, self.arena.slot_of):
            self.state[p]["step"] = torch.tensor(float(st.step[slot]))
        return st

    def state_dict(self):
        st = self._sync_steps()
        sd = super().state_dict()
        # tensors that never received a gradient have no entry in torch's state (their grad stays None)
        for idx, (p, slot) in enumerate(zip(self.arena.params, self.arena.slot_of)):
            if st.step[slot] == 0 and idx in sd["state"]:
                del sd["state"][idx]
        for k, v in sd["state"].items():
            sd["state"][k] = {kk: (vv.clone() if torch.is_tensor(vv) else vv) for kk, vv in v.items()}
        return sd

    def load_state_dict(self, state_dict):
        steps = [0] * 8
        for idx, (p, slot) in enumerate(zip(self.arena.params, self.arena.slot_of)):
            ent = state_dict["state"].get(idx)
            if ent is None:
                continue
            self.arena.view(self.arena.m, idx).copy_(ent["exp_avg"])
            self.arena.view(self.arena.v, idx).copy_(ent["exp_avg_sq"])
            steps[slot] = max(steps[slot], int(float(ent["step"])))
        self.arena.write_steps(steps)
        for g, sg in zip(self.param_groups, state_dict["param_groups"]):
            for k in ("lr", "betas", "eps", "weight_decay"):
                if k in sg:
                    g[k] = sg[k]

    def sync_lr_to_device(self):
        """Call before replaying a captured step: publishes param_groups[0]['lr'] to the device-side optimizer state."""
        self.arena.set_device_lr(float(self.param_groups[0]["lr"]))

    def last_grad_norm(self) -> float:
        return float(self.arena.read_state().last_gradnorm)

    def skipped_steps(self) -> int:
        return int(self.arena.read_state().skipped)

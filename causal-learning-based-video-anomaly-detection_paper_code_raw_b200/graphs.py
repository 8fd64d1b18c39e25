"""CUDA-graph capture of a whole train / inference step (forward, fused loss, backward, all-reduce, fused clip+Adam).

A step of the hot path is 250-300 kernel launches of a few microseconds each; issued one by one from Python the host
is slower than the device (~17 ms of launch work for a ~10 ms step).  Every kernel of libcvad_b200.so launches on the
caller's stream, allocates nothing and never synchronises, and all step bookkeeping (NaN flags, skip logic, Adam step
counts) lives on the device, so the whole step can be captured once and replayed with one launch.

Capturing needs a few warm-up executions on a side stream (lazy one-time initialisation: function attributes, driver
entry points, torch's allocator).  Those executions are real steps, so every tensor a step mutates -- parameters, Adam
moments and counters, BatchNorm running statistics -- is snapshotted before and restored after them: building the graph
does not advance training.
"""
from __future__ import annotations

import torch

from . import ops


class GraphedStep:
    """``step_fn(*static_inputs) -> tuple of tensors`` captured into one CUDA graph.

    ``__call__(*inputs)`` copies the inputs into the static buffers (an H2D copy when they live in pinned host memory),
    replays the graph and returns the static output tensors (valid until the next replay)."""

    def __init__(self, step_fn, example_inputs, mutated=(), warmup: int = 3, pre_replay=None, between=None, tail_fn=None):
        """``between`` / ``tail_fn`` split the step in two graphs around an eagerly issued call: graph(step_fn) ->
        between() -> graph(tail_fn).  The data-parallel trainer uses it to keep the NCCL gradient all-reduce OUT of the
        captured graphs (``between``) while forward/backward and the fused optimizer are each one graph launch."""
        self.pre_replay = pre_replay
        self.between, self.tail_graph = between, None
        self.static_inputs = [torch.empty_like(t, device=t.device if t.is_cuda else torch.cuda.current_device()) for t in example_inputs]
        for s, t in zip(self.static_inputs, example_inputs):
            s.copy_(t)
        mutated = [t for t in mutated if t is not None]
        snap = [t.detach().clone() for t in mutated]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn(*self.static_inputs)
                if between is not None:
                    between()
                if tail_fn is not None:
                    tail_fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            for t, s in zip(mutated, snap):
                t.copy_(s)
        del snap
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCHES[0]
        with torch.cuda.graph(self.graph):
            out = step_fn(*self.static_inputs)
        if tail_fn is not None:
            self.tail_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.tail_graph):
                tail_fn()
        self.launches = ops.LAUNCHES[0] - n0          # cvad ABI calls inside one replay
        self.outputs = out if isinstance(out, tuple) else (out,)

    def _replay(self):
        if self.pre_replay is not None:
            self.pre_replay()
        self.graph.replay()
        if self.between is not None:
            self.between()
        if self.tail_graph is not None:
            self.tail_graph.replay()
        ops.LAUNCHES[0] += self.launches

    # ---- input prefetch: the H2D copy of batch i+1 runs on a copy stream while the graph of batch i executes
    def _init_prefetch(self):
        self.staging = [torch.empty_like(s) for s in self.static_inputs]
        self.copy_stream = torch.cuda.Stream()
        self.ev_ready, self.ev_free = torch.cuda.Event(), torch.cuda.Event()
        self.ev_free.record(torch.cuda.current_stream())

    def prefetch(self, *inputs):
        """Start copying the NEXT step's inputs (pinned host or device tensors) into a staging buffer, asynchronously."""
        if not hasattr(self, "staging"):
            self._init_prefetch()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_free)           # the previous step has consumed the staging buffer
            for st, t in zip(self.staging, inputs):
                st.copy_(t, non_blocking=True)
            self.ev_ready.record(self.copy_stream)

    def run_prefetched(self):
        """Run one step on the inputs handed to the last ``prefetch`` call."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ev_ready)
        for s, st in zip(self.static_inputs, self.staging):
            s.copy_(st, non_blocking=True)                      # device-to-device, ~55 us for the 177 MB batch
        self.ev_free.record(cur)
        self._replay()
        return self.outputs

    def __call__(self, *inputs):
        for s, t in zip(self.static_inputs, inputs):
            if t is not s:
                s.copy_(t, non_blocking=True)
        self._replay()
        return self.outputs


def graphed_optimizer_step(opt, fwd_bwd, example_inputs, mutated):
    """Capture ``fwd_bwd(*inputs) -> tuple`` (zero_grad + forward + loss + backward) followed by ``opt``'s fused step.

    Single GPU: one graph.  Data parallel (the optimizer carries a gradient all-reduce hook): graph(fwd_bwd) -> eager collective ->
    graph(fused clip + Adam), because NCCL captured inside a graph stalled an 8-rank run."""
    import os

    opt.sync_lr_to_device()
    if opt.pre_step_hook is not None and os.environ.get("CVAD_NCCL_IN_GRAPH", "0") != "1":
        return GraphedStep(fwd_bwd, example_inputs, mutated, pre_replay=opt.sync_lr_to_device,
                           between=lambda: opt.pre_step_hook(opt.arena), tail_fn=opt.step_local)

    def step(*inputs):
        out = fwd_bwd(*inputs)
        opt.step()
        return out
    return GraphedStep(step, example_inputs, mutated, pre_replay=opt.sync_lr_to_device)

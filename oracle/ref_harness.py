"""Import harness for the UNMODIFIED reference scripts.

TEST INFRASTRUCTURE ONLY.  Two users:
  * tools/make_golden.py (build container) imports the scripts from /root/reference to (a) pin the oracle restatement in
    oracle/ against the real reference code and (b) emit the golden fixtures committed under tests/golden/;
  * bench.py --impl reference / its cpu_baseline leg import them from ``oracle/_ref`` -- a git-ignored, verbatim copy of the
    reference's scripts made by ``oracle/build_ref.py`` (run by ``__graft_entry__.build()`` while /root/reference is visible); the
    copy travels to the GPU box with the snapshot so the timed CPU arm is the reference's own code, Python loops included.
Nothing in the product imports it.

Shims:
  * stub ``matplotlib`` / ``seaborn`` / ``yolov5`` modules (not installed here);
  * ``ReduceLROnPlateau(verbose=...)`` raises on torch 2.11 -> subclass that drops it;
  * cad / cad1 call ``torch.manual_seed(42)`` at import -> callers reseed afterwards;
  * RNG injection: ``torch.rand_like`` / ``torch.randn_like`` and ``nn.Dropout`` are
    patched so the reference consumes caller-supplied noise / keep-masks.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import sys
import types

import torch
import torch.nn as nn

import os

REF_DIR = "/root/reference"
LOCAL_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_dir() -> str | None:
    """Where the unmodified reference scripts can be imported from: /root/reference (build container) or oracle/_ref (GPU box)."""
    for d in (REF_DIR, LOCAL_REF_DIR):
        if os.path.exists(os.path.join(d, "causal_anomaly_detection.py")):
            return d
    return None


class _Anything:
    """Attribute sink: any attribute access / call returns another sink."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


def _stub_module(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    def _ga(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything()

    m.__getattr__ = _ga  # type: ignore[attr-defined]
    m.__path__ = []  # behave like a package so "import a.b" works
    return m


def install_stubs() -> None:
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.animation",
                 "matplotlib.gridspec", "matplotlib.colors", "seaborn", "yolov5"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _stub_module(name)
    # ReduceLROnPlateau(verbose=True) shim
    sched = torch.optim.lr_scheduler
    if not getattr(sched.ReduceLROnPlateau, "_cvad_shim", False):
        base = sched.ReduceLROnPlateau

        class ReduceLROnPlateau(base):  # type: ignore[misc,valid-type]
            _cvad_shim = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)

        sched.ReduceLROnPlateau = ReduceLROnPlateau


def import_ref(modname: str, ref_dir: str | None = None):
    """Import one reference script as a module (its prints are swallowed)."""
    install_stubs()
    ref_dir = ref_dir or reference_dir() or REF_DIR
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    with contextlib.redirect_stdout(io.StringIO()):
        mod = importlib.import_module(modname)
    return mod


class NoiseInjector:
    """Feeds pre-drawn noise to the reference.

    ``rand`` / ``randn`` are FIFO lists of tensors returned by successive
    ``torch.rand_like`` / ``torch.randn_like`` calls.  ``dropout`` maps
    ``id(module)`` -> list of keep-masks (float 0/1); a Dropout in train mode
    computes ``x * mask / (1 - p)`` exactly like torch does with that mask.
    """

    def __init__(self):
        self.rand: list[torch.Tensor] = []
        self.randn: list[torch.Tensor] = []
        self.dropout: dict[int, list[torch.Tensor]] = {}
        self._orig = None

    def __enter__(self):
        self._orig = (torch.rand_like, torch.randn_like, nn.Dropout.forward)
        inj = self

        def rand_like(t, *a, **k):
            v = inj.rand.pop(0)
            assert v.shape == t.shape, (v.shape, t.shape)
            return v.to(t.dtype)

        def randn_like(t, *a, **k):
            v = inj.randn.pop(0)
            assert v.shape == t.shape, (v.shape, t.shape)
            return v.to(t.dtype)

        def drop_forward(mod, x):
            if not mod.training or mod.p == 0.0:
                return x
            q = inj.dropout.get(id(mod))
            assert q, "no dropout mask queued for this module"
            m = q.pop(0)
            assert m.shape == x.shape, (m.shape, x.shape)
            return x * m.to(x.dtype) * (1.0 / (1.0 - mod.p))

        torch.rand_like = rand_like
        torch.randn_like = randn_like
        nn.Dropout.forward = drop_forward
        return self

    def __exit__(self, *exc):
        torch.rand_like, torch.randn_like, nn.Dropout.forward = self._orig
        return False

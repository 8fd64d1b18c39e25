#!/usr/bin/env python
"""Sliding-window inference of M-A over a frame stream (16-frame windows every 4 frames, 240x360): frames/s of
ma0.StreamingWindowScorer (each frame's backbone pass computed once, windows gathered from a device ring) against scoring every
window as its own clip (the reference's loop shape, bbox:392-430 / cad test_model), same model, same GPU, bf16 backbone.

    python tools/stream_probe.py [frames] > profiles/rXX_streaming_windows.md
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cvad_b200  # noqa: E402,F401
from cvad_b200.ma import CausalAnomalyDetector  # noqa: E402
from cvad_b200.ma0 import StreamingWindowScorer  # noqa: E402
from cvad_b200.noise import FixedNoise  # noqa: E402
from test_oracle_golden import ma_synth_state  # noqa: E402


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    T, S, CH = 16, 4, int(os.environ.get('CHUNK', '512'))     # window, stride, frames pushed per call (= CHUNK/4 windows per call)
    dev = torch.device("cuda:0")
    m = CausalAnomalyDetector()
    m.load_state_dict(ma_synth_state(3, False), strict=True)
    m = m.to(dev).eval().set_precision("bf16")
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (F, 1, 240, 360), generator=g, dtype=torch.uint8).to(dev)
    n_win = (F - T) // S + 1
    eps = torch.zeros(n_win, 5, 6, device=dev)

    def stream():
        sc = StreamingWindowScorer(m, clip_len=T, stride=S, capacity=2 * CH + T)
        out, done = [], 0
        for p in range(0, F, CH):
            k = (max(p + CH - T, -1) // S + 1 if p + CH >= T else 0) - done
            if k > 0:
                m.noise = FixedNoise({"eps": eps[done:done + k]})
            s, _ = sc.push(frames[p:p + CH])
            out.append(s)
            done += s.shape[0]
        return torch.cat(out)

    def clipwise(batch=CH // S):
        out = []
        with torch.no_grad():
            for w0 in range(0, n_win, batch):
                ws = range(w0, min(w0 + batch, n_win))
                clips = torch.stack([frames[w * S: w * S + T] for w in ws])
                m.noise = FixedNoise({"eps": eps[w0:w0 + len(ws)]})
                out.append(m(clips)["anomaly_scores"])
        return torch.cat(out)

    res = {}
    for name, fn in (("streaming (backbone once per frame)", stream), ("window by window", clipwise)):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s = fn()
        torch.cuda.synchronize()
        res[name] = (time.perf_counter() - t0, s)
    a, b = res["streaming (backbone once per frame)"][1], res["window by window"][1]
    err = float((a - b).abs().max() / b.abs().max())
    print(f"# M-A sliding-window inference over a stream of {F} frames (240x360 uint8), {T}-frame windows every {S} frames = {n_win} windows, {CH} frames = {CH // S} windows per call, bf16 backbone, eager launches (no CUDA graph), one B200\n")
    print("| path | seconds | stream frames/s | windows/s | backbone frame passes |\n|---|---:|---:|---:|---:|")
    for name, (dt, _) in res.items():
        passes = F if name.startswith("streaming") else n_win * T
        print(f"| {name} | {dt:.4f} | {F / dt:,.0f} | {n_win / dt:,.0f} | {passes:,} |")
    print(f"\nmax relative difference of the window scores between the two paths: {err:.2e} (eval-mode BatchNorm: a frame's features do not depend on its window)")


if __name__ == "__main__":
    main()

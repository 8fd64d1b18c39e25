#!/usr/bin/env python
"""Achieved HBM bandwidth of the bandwidth-bound kernel families INSIDE the replayed benchmark step: algorithmic bytes at the benchmark shapes
(512 frames of 240x360; every operand read once, every result written once) / the per-call time of `bench.py --profile-calls`.

    python tools/bandwidth_in_step.py profiles/rXX_calls_in_graph.md > profiles/rXX_bandwidth_in_step.md
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 512
LAYERS = [(60, 90, 32, False), (60, 90, 32, True), (30, 45, 64, False), (30, 45, 64, True), (15, 23, 128, False), (15, 23, 128, True),
          (8, 12, 256, False), (8, 12, 256, False)]          # (H, W, C, output written as phase planes) of the eight BatchNorm layers


def padded(h, w, c, phase):
    if not phase:
        return N * (h + 2) * (w + 2) * c * 2
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    return 4 * N * (ho + 2) * (wo + 2) * c * 2


def main():
    calls = {}
    for line in open(sys.argv[1]):
        m = re.match(r"\| `(\w+)(?:\[.*\])?` \| (\d+) \| ([\d.]+) \|", line)
        if m:
            calls[m.group(1)] = calls.get(m.group(1), 0.0) + float(m.group(3))
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6499.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    interior = [N * h * w * c * 2 for h, w, c, _ in LAYERS]
    fam = [
        ("cvad_pad_bn_apply_relu_bf16", "BatchNorm + ReLU apply, 8 layers (read raw interior, write act incl. zero border / phase planes)",
         sum(i + padded(h, w, c, ph) for i, (h, w, c, ph) in zip(interior, LAYERS))),
        ("cvad_pad_bn_relu_bwd_bf16", "ReLU + BatchNorm backward, 8 layers (reduce: raw + dact; apply: raw + dact -> draw)",
         sum(4 * i + padded(h, w, c, False) for i, (h, w, c, _) in zip(interior, LAYERS))),
        ("cvad_stem_space_to_depth_u8", "uint8 frames -> normalised fp32 2x2 space-to-depth", N * 240 * 360 + N * 123 * 183 * 16),
        ("cvad_pad_avgpool_bf16_fwd", "AdaptiveAvgPool(4,6) forward", N * 8 * 12 * 256 * 2 + N * 6144 * 4),
        ("cvad_pad_avgpool_bf16_bwd", "AdaptiveAvgPool(4,6) backward", N * 8 * 12 * 256 * 2 + N * 6144 * 4),
        ("cvad_adam_flat_f32", "fused clip + AdamW over the flat arena (p, g, m, v read; p, m, v written)", 7938048 * 4 * 7),
        ("cvad_sumsq_f32", "gradient norm^2 + non-finite test over the gradient arena", 7938048 * 4),
    ]
    print(f"# bandwidth-bound kernel families inside the replayed M-A train step (batch 32 = 512 frames; source: {os.path.basename(sys.argv[1])})\n")
    print(f"peak = {peak:.0f} GB/s (MEASURED_PEAKS.json copy bandwidth).  Times are external-CUDA-event brackets inside the instrumented graph, which add ~3 us per call:")
    print("the fractions of the short calls are lower bounds.\n")
    print("| ABI call | what | MB | us | GB/s | frac of peak |\n|---|---|---:|---:|---:|---:|")
    tb = tu = 0.0
    for name, what, nbytes in fam:
        if name not in calls:
            continue
        us = calls[name]
        gbs = nbytes / us / 1e3
        print(f"| `{name}` | {what} | {nbytes / 1e6:.0f} | {us:.1f} | {gbs:.0f} | {gbs / peak:.2f} |")
        if "bn_" in name:
            tb += nbytes
            tu += us
    if tu:
        print(f"\nBatchNorm family in the step: {tb / 1e6:.0f} MB in {tu:.0f} us = {tb / tu / 1e3:.0f} GB/s ({tb / tu / 1e3 / peak:.2f} of peak).")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Achieved HBM bandwidth of the bandwidth-bound kernels of the hot path at the benchmark shapes (512 frames of M-A, batch 32):
BatchNorm statistics / apply / backward over the padded-flat bf16 activations, max-pool, avg-pool, the fused clip+AdamW step,
the M-D reconstruction loss and the M-B 5-term loss.  CUDA events, L2 flushed before every timed launch, algorithmic bytes
(each operand read once, each result written once) / time against the measured copy bandwidth (MEASURED_PEAKS.json).

    python tools/bw_probe.py [frames] > profiles/rXX_bandwidth_kernels.md
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cvad_b200  # noqa: E402,F401
from cvad_b200 import ops, tc  # noqa: E402
from cvad_b200.ops import _call, _ptr, _st  # noqa: E402

LAYERS = [(60, 90, 32, False), (60, 90, 32, True), (30, 45, 64, False), (30, 45, 64, True), (15, 23, 128, False), (15, 23, 128, True),
          (8, 12, 256, False), (8, 12, 256, False)]          # (H, W, C, output written as phase planes)


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)).get("hbm_gbs", 6499.0) if os.path.exists(p) else 6650.0


def timed(fn, flush, reps=5):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda:0")
    bf = torch.bfloat16
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    pk = peak()
    rows = []

    def add(name, shape, nbytes, ms):
        gbs = nbytes / ms / 1e6
        rows.append((name, shape, nbytes / 1e6, ms * 1e3, gbs, gbs / pk))

    st = _st()
    for (H, W, C, phase) in LAYERS:
        raw = (torch.randn(N, H + 2, W + 2, C, device=dev) * 2).to(bf)
        mean, invstd = torch.empty(C, device=dev), torch.empty(C, device=dev)
        gam, bet = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
        rm, rv, nbt = torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.tensor(0, device=dev)
        ws = ops.bn_workspace(dev, C)
        interior = N * H * W * C * 2
        shape = f"{N}x{H}x{W}x{C}" + (" ->phase" if phase else "")
        add("pad_bn_stats (sum, sumsq)", shape, interior,
            timed(lambda: _call("cvad_pad_bn_stats_bf16", _ptr(raw), N, H, W, C, _ptr(ws), 1e-5, 0.1, _ptr(mean), _ptr(invstd), _ptr(rm), _ptr(rv),
                                _ptr(nbt), st), flush))
        act = torch.empty(tc.act_shape(N, H, W, C, phase), device=dev, dtype=bf)
        add("pad_bn_apply_relu", shape, interior + act.numel() * 2,
            timed(lambda: _call("cvad_pad_bn_apply_relu_bf16", _ptr(raw), _ptr(act), N, H, W, C, int(phase), _ptr(mean), _ptr(invstd), _ptr(gam),
                                _ptr(bet), st), flush))
        dact = torch.randn(tc.act_shape(N, H, W, C, phase), device=dev).to(bf)
        draw = torch.empty_like(raw)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        # reduce pass reads raw + dact, apply pass reads raw + dact and writes draw (complete buffer)
        add("pad_bn_relu_bwd (reduce + apply)", shape, 4 * interior + draw.numel() * 2,
            timed(lambda: _call("cvad_pad_bn_relu_bwd_bf16", _ptr(raw), _ptr(dact), _ptr(draw), N, H, W, C, int(phase), _ptr(mean), _ptr(invstd),
                                _ptr(gam), _ptr(bet), 1, _ptr(ws), _ptr(dg), _ptr(db), st), flush))
        del raw, act, dact, draw
    # stem max-pool 3x3 s2 (120x180x32 -> 60x90x32 padded-flat)
    y1 = torch.randn(N, 120, 180, 32, device=dev).to(bf)
    a0 = torch.empty(N, 62, 92, 32, device=dev, dtype=bf)
    add("maxpool3x3s2 -> padded-flat", f"{N}x120x180x32", y1.numel() * 2 + a0.numel() * 2,
        timed(lambda: _call("cvad_pad_maxpool3x3s2_bf16", _ptr(y1), N, 120, 180, 32, _ptr(a0), st), flush))
    del y1, a0
    # adaptive avg-pool (8x12x256 -> 4x6) forward / backward
    a8 = torch.randn(N, 10, 14, 256, device=dev).to(bf)
    feats = torch.empty(N, 256, 4, 6, device=dev)
    add("avgpool (4,6) fwd", f"{N}x8x12x256", N * 8 * 12 * 256 * 2 + feats.numel() * 4,
        timed(lambda: _call("cvad_pad_avgpool_bf16_fwd", _ptr(a8), N, 8, 12, 256, 4, 6, _ptr(feats), st), flush))
    add("avgpool (4,6) bwd", f"{N}x8x12x256", a8.numel() * 2 + feats.numel() * 4,
        timed(lambda: _call("cvad_pad_avgpool_bf16_bwd", _ptr(feats), N, 8, 12, 256, 4, 6, _ptr(a8), st), flush))
    del a8, feats
    # fused grad-norm + clip + AdamW over the M-A arena (trainable parameters; p, g, m, v read, p, m, v written)
    from cvad_b200.ma import CausalAnomalyDetector, MATrainer
    tr = MATrainer(CausalAnomalyDetector(), dev, precision="bf16")
    ar = tr.optimizer.arena
    ar.g.normal_()
    ar.g[:16] = 0
    ar.g[1:8] = 1.0
    add("sumsq + clip + AdamW (flat arena)", f"{ar.total} fp32", ar.total * 4 * (1 + 4 + 3), timed(lambda: tr.optimizer.step_local(), flush))
    del tr, ar
    # M-D reconstruction MSE (cad1:323-344) at 256 clips x 8 frames x 64x64
    B, T, E = 256, 8, 4096
    recon, frames = torch.rand(B, T, E, device=dev), torch.rand(B, T, E, device=dev)
    dr, clip, loss = torch.empty_like(recon), torch.empty(B, device=dev), torch.empty(1, device=dev)
    wsd = torch.zeros(B, device=dev, dtype=torch.float64)
    add("recon MSE + gradient (M-D, cad1:323-344)", f"{B}x{T}x{E} fp32", B * T * E * 4 * 3,
        timed(lambda: _call("cvad_recon_mse_f32", _ptr(recon), E, _ptr(frames), B, T, E, _ptr(wsd), _ptr(clip), _ptr(loss), _ptr(dr), None, st),
              flush))
    # M-B fused 5-term loss + gradients (s2:135-205) at the benchmark batch of 32 clips: latency-bound (the pair term is O(B^2))
    Bm = 32
    sc, adj = torch.rand(Bm, device=dev), torch.rand(Bm, 256, device=dev)
    pseudo = (torch.rand(Bm, device=dev) > 0.95).float()
    wsf = torch.empty(int(ops.L().cvad_mb_loss_ws_floats(Bm)), device=dev)
    out8, ds, da = torch.empty(8, device=dev), torch.empty(Bm, device=dev), torch.empty(Bm, 256, device=dev)
    add("M-B 5-term loss + gradients (latency-bound)", f"{Bm} clips", Bm * (257 * 4 * 2 + 4),
        timed(lambda: _call("cvad_mb_loss_f32", _ptr(sc), _ptr(adj), _ptr(pseudo), Bm, 1.0, 0.01, 0.001, 0.01, _ptr(wsf), _ptr(out8), _ptr(ds),
                            _ptr(da), None, st), flush))
    print(f"# bandwidth-bound kernels at the benchmark shapes ({N} frames), CUDA events, L2 flushed before each launch\n")
    print(f"peak = {pk:.0f} GB/s (MEASURED_PEAKS.json copy bandwidth); bytes = algorithmic (each operand once)\n")
    print("| kernel | shape | MB | us | GB/s | frac of peak |\n|---|---|---:|---:|---:|---:|")
    for name, shape, mb, us, gbs, fr in rows:
        print(f"| `{name}` | {shape} | {mb:.1f} | {us:.1f} | {gbs:.0f} | {fr:.2f} |")
    tot_b = sum(r[2] for r in rows[:24])
    tot_us = sum(r[3] for r in rows[:24])
    print(f"\nBatchNorm family over the 8 layers: {tot_b:.0f} MB in {tot_us:.0f} us = {tot_b / tot_us * 1e3:.0f} GB/s "
          f"({tot_b / tot_us * 1e3 / pk:.2f} of peak)")


if __name__ == "__main__":
    main()

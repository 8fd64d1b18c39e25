#!/usr/bin/env python
"""Headline benchmark: M-A (causal_anomaly_detection.py) training step on synthetic Avenue-shaped clips.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ma_train|mb_train|mc_infer]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch: zero_grad -> forward ->
4-term loss -> backward -> (gradient all-reduce when N > 1) -> fused clip + AdamW.
  value     clips/s with inputs resident in HBM (the 177 MB fp32 batch exceeds the 126 MB L2, so every step streams it)
  e2e       the same through the public trainer API with the batch in PINNED HOST memory: H2D copy of the step's inputs and
            a D2H read of the loss inside the timed region
  roofline  the dominant kernel family (the tcgen05 implicit-GEMM convolutions) timed with CUDA events inside the timed
            region against the measured dense bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle port of the reference step on the box's host cores, bounded sample (rank 0, N=1 only)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

T, H, W = 16, 240, 360
PER_GPU_BATCH = 32
# algorithmic FLOPs (2*MAC) per frame of the eight 3x3 convolutions: forward, and forward+dgrad+wgrad
CONV_FWD_MFLOP = [99.53, 99.53, 49.77, 99.53, 50.87, 101.74, 56.62, 113.25]
CONV_TRAIN_MFLOP = sum(CONV_FWD_MFLOP) * 3 - CONV_FWD_MFLOP[0]      # layer1.0 needs no data-gradient (frozen stem below it)
# compulsory HBM bytes of the same 23 calls for 512 frames: every operand read once and every result written once in bf16 over the
# logical (unpadded) tensors; weights and weight-gradients are negligible.  (elements per frame of each layer's input / output)
_CONV_IO = [(172800, 172800), (172800, 172800), (172800, 86400), (86400, 86400), (86400, 44160), (44160, 44160), (44160, 24576), (24576, 24576)]
CONV_ALG_BYTES = sum(2 * 512 * ((i + o) + (i + o) + ((i + o) if k else 0)) for k, (i, o) in enumerate(_CONV_IO))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1341.2), d.get("hbm_gbs", 6499.0), "measured"
    return 1400.0, 6650.0, "fallback"


def conv_family_traffic():
    """DRAM bytes the convolution family moves per 512-frame step, from the committed ncu capture (tools/conv_traffic.py)."""
    p = os.path.join(ROOT, "profiles", "conv_family_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    return d.get("bytes_per_step"), d.get("source")


_T0 = time.time()


def trace(msg):
    """Stage markers on stderr (CVAD_BENCH_TRACE=1): where every rank is, should a multi-GPU run stall."""
    if os.environ.get("CVAD_BENCH_TRACE", "0") == "1":
        print(f"[bench rank {os.environ.get('RANK', '0')} +{time.time() - _T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons}


def synth_batch(B, seed):
    import synth
    x = synth.ma_clips(B, T, H, W, seed, wide=True)
    y = (torch.rand(B, generator=synth.gen(seed + 9)) < 0.3).long()
    return x, y


def cpu_reference_step_rate(sample_clips=2, steps=2, warm=1):
    """Oracle port of the reference train step on the host cores: clips/s on a bounded sample."""
    from oracle import train as o_train
    from test_oracle_golden import ma_synth_state
    import synth
    torch.set_num_threads(os.cpu_count() or 1)
    P = ma_synth_state(3, False)
    opt = o_train.OracleAdam(o_train.ma_trainable(P), 3e-4, 1e-5, True, 1.0)
    x, y = synth_batch(sample_clips, 1234)
    eps = torch.randn(sample_clips, 5, 6, generator=synth.gen(1))
    keep = {"det0": synth.keep_mask((sample_clips, T, 512), 0.3, 2), "det1": synth.keep_mask((sample_clips, T, 256), 0.2, 3),
            "scorer0": synth.keep_mask((sample_clips, 64), 0.2, 4), "cls0": synth.keep_mask((sample_clips, 512), 0.3, 5),
            "cls1": synth.keep_mask((sample_clips, 256), 0.2, 6)}
    for _ in range(warm):
        o_train.ma_train_step(P, opt, x, y, eps, keep)
    t0 = time.perf_counter()
    for _ in range(steps):
        o_train.ma_train_step(P, opt, x, y, eps, keep)
    dt = (time.perf_counter() - t0) / steps
    return sample_clips / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 2
    rate, dt, cores = cpu_reference_step_rate(sample, max(1, min(args.steps, 3)), max(1, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": "train clips/sec", "value": rate, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "M-A (causal_anomaly_detection.py) train step, synthetic Avenue clips (B,16,1,240,360)",
                   "per_step_sample_clips": sample},
        "cpu_baseline": {"value": rate, "unit": "clips/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} clips/step (fwd + 4-term loss + bwd + clip + AdamW), oracle port of cad:637-690 on CPU"},
        "e2e": {"value": rate, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def profile_calls(tr, x_dev, y_dev, path, reps=5):
    """Per-ABI-call time inside the replayed step graph (external CUDA events around EVERY call of libcvad_b200.so), written as a
    markdown table: where the step's milliseconds go, including what is not ours (torch fills / RNG) as the unbracketed rest."""
    from cvad_b200 import ops
    ops.TIMED.clear()
    ops.TIMED_NAMES.add("*")
    ops.TIMED_CAPTURE_ONLY[0] = True
    gp = tr.graphed_train_step(x_dev, y_dev)
    ops.TIMED_CAPTURE_ONLY[0] = False
    ops.TIMED_NAMES.clear()
    acc = {k: 0.0 for k in ops.TIMED}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for i in range(reps + 1):
        e0.record()
        gp(x_dev, y_dev)
        e1.record()
        torch.cuda.synchronize()
        if i:
            total += e0.elapsed_time(e1) / reps
            for k, v in ops.TIMED.items():
                acc[k] += sum(s.elapsed_time(e) for s, e in v) / reps
    lines = [f"# per-call time inside the replayed M-A train-step graph (batch 32, external CUDA events, mean of {reps} replays)\n",
             f"step (instrumented graph) {total * 1e3:.0f} us; sum of bracketed calls {sum(acc.values()) * 1e3:.0f} us "
             "(calls on the side streams overlap the main chain, so the sum may exceed the step)\n",
             "| ABI call | calls | us | share of step |", "|---|---:|---:|---:|"]
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        lines.append(f"| `{k}` | {len(ops.TIMED[k])} | {v * 1e3:.1f} | {100 * v / total:.1f}% |")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    ops.TIMED.clear()


def run_ours(args):
    # NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION is set in the image; stdout carries exactly one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch.distributed as dist
    import cvad_b200
    from cvad_b200 import ops
    from cvad_b200.ma import CausalAnomalyDetector, MATrainer
    from cvad_b200.parallel import DataParallel, init_from_env

    trace("imports done")
    rank, local, world = init_from_env()
    trace(f"process group up (world {world})")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    torch.manual_seed(1234 + rank)
    B = PER_GPU_BATCH
    dp = DataParallel() if world > 1 else None
    tr = MATrainer(CausalAnomalyDetector(), dev, precision=args.precision, dp=dp)
    if dp is not None:
        dp.broadcast_parameters(tr.optimizer.arena)
        torch.cuda.synchronize()
        trace("parameters broadcast")
    tr.model.train()
    x_host, y_host = synth_batch(B, 1234 + rank)
    x_pin, y_pin = x_host.pin_memory(), y_host.pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.no_graph:
        launches_per_step = None

        def step(x, y):
            comp, _ = tr.train_step(x, y)
            return comp
    else:
        # the whole step (zero_grad, forward, loss, backward, all-reduce, clip+AdamW) is one CUDA graph
        trace("batch resident; capturing the step")
        gs = tr.graphed_train_step(x_dev, y_dev)
        trace("step captured")
        launches_per_step = gs.launches
        x_dev, y_dev = gs.static_inputs

        def step(x, y):
            return gs(x, y)[0]

    for _ in range(args.warmup):
        step(x_dev, y_dev)
    barrier()
    trace("warm-up done")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = ops.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(x_dev, y_dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    trace(f"timed region done ({ms / args.steps:.3f} ms/step)")
    launches = ops.LAUNCHES[0] - n0
    # ---- end to end: pinned host batch -> H2D -> step -> D2H loss, through the public trainer API
    # Every step's batch starts in pinned host memory and its loss is read back to the host.  With the graphed trainer the
    # H2D copy of step i+1 is issued before step i's loss is read, so it overlaps step i's compute (all copies are inside
    # the timed region: K+1 batches are copied for K timed steps).
    def e2e_step(first):
        if args.no_graph:
            return step(x_pin.to(dev, non_blocking=True), y_pin.to(dev, non_blocking=True))
        if first:
            gs.prefetch(x_pin, y_pin)
        out = gs.run_prefetched()[0]
        gs.prefetch(x_pin, y_pin)
        return out

    for i in range(2):
        float(e2e_step(i == 0)[0])
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        c = e2e_step(False)
        loss_host = float(c[0])
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    trace(f"e2e region done ({ms_e2e / args.steps:.3f} ms/step)")
    sampler.stop_flag = True
    # ---- dominant kernel family, timed live with CUDA events around every call (eager launches of the same step)
    conv_names = {"cvad_flat_conv3x3_fwd_bf16", "cvad_flat_conv3x3_fwd_stats_bf16", "cvad_flat_conv3x3_dgrad_bf16",
                  "cvad_flat_conv3x3_wgrad_bf16", "cvad_flat_conv3x3_wgrad_staged_bf16"} if args.precision == "bf16" \
        else {"cvad_conv_fwd_f32", "cvad_conv_dgrad_f32", "cvad_conv_wgrad_f32"}
    probe_steps = 5
    ops.TIMED.clear()
    ops.TIMED_NAMES.update(conv_names)
    if args.no_graph:
        for _ in range(2):
            tr.train_step(x_dev, y_dev)
        torch.cuda.synchronize()
        ops.TIMED.clear()
        for _ in range(probe_steps):
            tr.train_step(x_dev, y_dev)
        torch.cuda.synchronize()
        conv_ms = {k: sum(s.elapsed_time(e) for s, e in v) / probe_steps for k, v in ops.TIMED.items()}
        how = f"CUDA events around each ABI call over {probe_steps} eager steps after the timed region"
    else:
        # a second capture of the same step with an (external) CUDA event pair around every convolution call: the events are
        # nodes of the graph, so each replay times the kernels exactly as they run inside the benchmarked step
        ops.TIMED_CAPTURE_ONLY[0] = True
        gp = tr.graphed_train_step(x_dev, y_dev)
        ops.TIMED_CAPTURE_ONLY[0] = False
        ops.TIMED_NAMES.clear()
        conv_ms = {k: 0.0 for k in ops.TIMED}
        for i in range(probe_steps + 1):
            gp(x_dev, y_dev)
            torch.cuda.synchronize()
            if i:                                   # the first replay warms the instrumented graph up
                for k, v in ops.TIMED.items():
                    conv_ms[k] += sum(s.elapsed_time(e) for s, e in v) / probe_steps
        how = f"external CUDA events around each ABI call inside the replayed step graph, mean of {probe_steps} replays after the timed region"
    ops.TIMED_NAMES.clear()
    conv_launches = sum(len(v) for v in ops.TIMED.values()) // (probe_steps if args.no_graph else 1)
    if args.profile_calls and not args.no_graph and rank == 0:
        profile_calls(tr, x_dev, y_dev, args.profile_calls)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    trace("kernel probe done")
    if rank != 0:
        return
    tf_peak, hbm_peak, src = peaks()
    traffic, traffic_src = conv_family_traffic() if args.precision == "bf16" else (None, None)
    ms_step = ms / args.steps
    value = world * B / (ms_step / 1e3)
    e2e_value = world * B / (ms_e2e / args.steps / 1e3)
    conv_total_ms = sum(conv_ms.values())
    flops_step = CONV_TRAIN_MFLOP * 1e6 * B * T
    achieved = flops_step / (conv_total_ms / 1e3) / 1e12 if conv_total_ms > 0 else 0.0
    line = {
        "metric": "train clips/sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "M-A (causal_anomaly_detection.py) train step, synthetic Avenue clips (32,16,1,240,360) per GPU",
                   "per_gpu_batch": B, "frames_per_clip": T, "frame": [H, W], "parallelism": f"dp{world}",
                   "l2": "inputs (177 MB fp32 per step) larger than the 126 MB L2",
                   "launch": "eager" if args.no_graph else "one CUDA graph per step",
                   "e2e_path": "pinned host batch -> H2D (copy stream, overlapped with the previous step) -> graph -> D2H loss",
                   "precision": "bf16 operands, fp32 accumulate (tcgen05), fp32 stem/tail" if args.precision == "bf16" else "fp32"},
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": x_pin.numel() * 4 + y_pin.numel() * 8, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak, "traffic": traffic,
                     "traffic_source": traffic_src, "algorithmic_bytes_per_step": CONV_ALG_BYTES,
                     "kernel": "flatconv / flatwgrad TMA+tcgen05 3x3 convolutions (fwd incl. BatchNorm statistics + dgrad + wgrad incl. its fold pass, 8 layers)" if args.precision == "bf16"
                     else "conv_gemm_kernel fp32",
                     "peak_source": src, "launches_per_step": conv_launches, "ms_per_step": conv_total_ms,
                     "share_of_step": conv_total_ms / ms_step, "per_kernel_ms": conv_ms,
                     "algorithmic_gflop_per_step": flops_step / 1e9,
                     "how": how},
        "clocks": sampler.summary(),
        "loss": loss_host,
    }
    if world == 1 and not args.no_cpu:
        rate, dt, cores = cpu_reference_step_rate(2, 2, 1)
        line["cpu_baseline"] = {"value": rate, "unit": "clips/s", "cores": cores, "kind": "port",
                                "sample": "2 clips/step x 2 steps of the same workload (oracle port of cad:637-690, fp32, all host threads)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying one CUDA graph")
    ap.add_argument("--profile-calls", default="", help="write a per-ABI-call timing table of the replayed step graph to this file")
    ap.add_argument("--watchdog", type=int, default=300, help="seconds after which a stuck run dumps its stack and exits non-zero")
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(args.watchdog, exit=True)     # a hang must never eat the GPU budget
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    # Leave without tearing NCCL down: the captured step graphs still reference the communicator, and destroying it first can
    # block.  Everything that matters has been printed and flushed.
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()

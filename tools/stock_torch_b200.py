"""Context numbers (NOT a target): stock eager PyTorch (cuDNN / cuBLAS) on the same B200 for the benchmarked step.

  1. the UNMODIFIED reference (oracle/_ref) driven through its own train_model on device 'cuda': fp16 autocast + GradScaler as cad:621-668
     does, Python loops and host syncs included;
  2. the reference's own ResNetBackbone class alone (97 % of the step's FLOPs), forward + backward under bf16 autocast (cuDNN, benchmark mode)
     -- the part of the step that dominates, without the reference's Python loops.

    python tools/stock_torch_b200.py > gpurun_out/stock_torch_b200.json
"""
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402
from test_oracle_golden import ma_synth_state  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cad = ref_harness.import_ref("causal_anomaly_detection")
    cad.device = dev
    B, T, H, W = 32, 16, 240, 360
    x = synth.ma_clips(B, T, H, W, 1234, wide=True)
    y = (torch.rand(B, generator=synth.gen(1243)) < 0.3).long()
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "batch": [B, T, 1, H, W]}
    quiet = contextlib.redirect_stdout(io.StringIO())
    # ---- 1. the reference as it is
    torch.manual_seed(0)
    model = cad.CausalAnomalyDetector()
    model.load_state_dict(ma_synth_state(3, False), strict=True)
    xd, yd = x.to(dev), y.to(dev)
    with quiet:
        cad.train_model(model, [(xd, yd)] * 2, [], num_epochs=1, lr=3e-4)
    torch.cuda.synchronize()
    n = 5
    t0 = time.perf_counter()
    with quiet:
        cad.train_model(model, [(xd, yd)] * n, [], num_epochs=1, lr=3e-4)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    out["reference_unmodified_cuda_fp16_autocast"] = {"ms_per_step": dt * 1e3, "clips_per_s": B / dt,
                                                      "what": "causal_anomaly_detection.train_model on device cuda (its own AMP path), batch resident on the GPU"}
    del model
    torch.cuda.empty_cache()
    # ---- 2. backbone only, bf16 autocast + channels_last, forward + backward
    torch.manual_seed(0)
    # (channels_last is not an option for the unmodified class: its forward ends in x.view(B, T, -1), which rejects NHWC strides)
    bb = cad.ResNetBackbone(input_channels=1, output_dim=256).to(dev).train()
    for name, p in bb.named_parameters():
        if name.startswith("conv1") or name.startswith("bn1"):
            p.requires_grad = False
    torch.backends.cudnn.benchmark = True

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            f = bb(xd)
        f.float().square().mean().backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out["reference_backbone_only_bf16_autocast"] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3,
                                                                  "what": "cad.ResNetBackbone forward + backward only (no detector / tail / loss / optimizer), cuDNN, frozen stem"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

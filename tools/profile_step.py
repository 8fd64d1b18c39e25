#!/usr/bin/env python
"""The benchmarked M-A train step (batch 32, bf16, uint8 frames) captured into its CUDA graph and replayed a few times -- nothing else.
Meant to run UNDER ncu: the launch list of the last replay is the launch list of one benchmarked step.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
    python tools/ncu_summary.py gpurun_out/launches.csv > profiles/rXX_launches_ma_train_b32.md
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from test_oracle_golden import ma_synth_state  # noqa: E402
from cvad_b200.ma import CausalAnomalyDetector, MATrainer  # noqa: E402


def main():
    replays = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda:0")
    model = CausalAnomalyDetector()
    model.load_state_dict(ma_synth_state(3, False), strict=True)
    tr = MATrainer(model, dev, precision="bf16")
    tr.model.train()
    torch.manual_seed(1234)
    x, y = bench.synth_batch(32, 1234, "u8")
    gs = tr.graphed_train_step(x.to(dev), y.to(dev))
    for _ in range(replays):
        out = gs(*gs.static_inputs)
    torch.cuda.synchronize()
    print("loss", float(out[0][0]))


if __name__ == "__main__":
    main()

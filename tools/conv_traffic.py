#!/usr/bin/env python
"""DRAM traffic of the convolution family per train step, from the ncu metrics pass over tools/conv_probe.py
(profiles/r01f_conv_kernels_ncu_metrics.md: dram__bytes_read.sum / dram__bytes_write.sum per launch, REPS=1, cold cache).

conv_probe launches, per layer, forward x2, data-gradient x2 (four phase launches each for the stride-2 layers) and
weight-gradient x2; the second (timed) set of each is taken.  A train step runs forward + weight-gradient of all eight layers
and the data-gradient of layers 1..7 (layer1.0 sits on the frozen stem).  Usage: python tools/conv_traffic.py > profiles/conv_family_traffic.json"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "profiles/r01f_conv_kernels_ncu_metrics.md"
STRIDES = [1, 1, 2, 1, 2, 1, 2, 1]


def main():
    rows = [l.strip().split("|") for l in open(os.path.join(ROOT, SRC)) if re.match(r"\| \d+ \|", l)]
    R = [(float(r[4]), float(r[5]), float(r[6])) for r in rows]          # us, read MB, write MB
    idx, per_layer, total_mb, launches = 0, [], 0.0, 0
    for li, s in enumerate(STRIDES):
        nd = 1 if s == 1 else 4
        f = R[idx + 1: idx + 2]; idx += 2
        d = R[idx + nd: idx + 2 * nd]; idx += 2 * nd
        w = R[idx + 1: idx + 2]; idx += 2
        mb = lambda xs: sum(x[1] + x[2] for x in xs)
        use = mb(f) + mb(w) + (mb(d) if li > 0 else 0.0)
        launches += 2 + (nd if li > 0 else 0)
        total_mb += use
        per_layer.append({"layer": li, "fwd_mb": round(mb(f), 1), "dgrad_mb": round(mb(d), 1), "wgrad_mb": round(mb(w), 1)})
    assert idx == len(R), (idx, len(R))
    json.dump({"source": SRC, "what": "dram__bytes_read.sum + dram__bytes_write.sum summed over the convolution launches of one 512-frame train step",
               "launches_per_step": launches, "bytes_per_step": int(total_mb * 1e6), "per_layer": per_layer}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""DRAM traffic of the convolution family per train step from an ncu metrics pass over `STEP_ONLY=1 python tools/conv_probe.py 512`
(which launches exactly the convolution kernels of one 512-frame step, once each):

    STEP_ONLY=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/conv_step_metrics.csv python tools/conv_probe.py 512
    python tools/conv_traffic.py gpurun_out/conv_step_metrics.csv profiles/rXX_conv_step_metrics.md > profiles/conv_family_traffic.json
"""
import collections
import csv
import json
import re
import sys


def main():
    path, md = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", r["Kernel Name"]).split("(")[0]
        if not any(t in k for t in ("flatconv_kernel", "flatwgrad_kernel", "wgrad_fold_kernel")):
            continue
        d = per.setdefault(r["ID"], {"kernel": k, "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        if "byte" in unit:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        elif unit in ("us", "usecond"):
            v *= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e6
        d[r["Metric Name"]] = v
    rows = list(per.values())
    total = sum(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"] for r in rows)
    t = sum(r["gpu__time_duration.sum"] for r in rows)
    if md:
        with open(md, "w") as f:
            f.write("# convolution launches of one 512-frame train step: ncu metrics pass (cold cache, serialised)\n\n")
            f.write("| # | kernel | grid | us | DRAM read MB | DRAM write MB |\n|---|---|---|---:|---:|---:|\n")
            for i, r in enumerate(rows):
                f.write(f"| {i} | `{r['kernel']}` | {r['grid']} | {r['gpu__time_duration.sum'] / 1e3:.1f} | {r['dram__bytes_read.sum'] / 1e6:.1f} | "
                        f"{r['dram__bytes_write.sum'] / 1e6:.1f} |\n")
            f.write(f"\n{len(rows)} launches, {t / 1e3:.0f} us, {total / 1e9:.3f} GB\n")
    json.dump({"source": md or path, "what": "dram__bytes_read.sum + dram__bytes_write.sum summed over the convolution launches of one 512-frame train step",
               "launches_per_step": len(rows), "bytes_per_step": int(total), "ncu_us_per_step": t / 1e3}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()

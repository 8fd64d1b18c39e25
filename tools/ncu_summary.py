#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches and device time of ONE train step
(the launches between two consecutive optimizer kernels).  Usage: python tools/ncu_summary.py launches.csv [step_index] > profiles/x.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else -2
    marker = sys.argv[3] if len(sys.argv) > 3 else "adam_flat"
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = [(x["Kernel Name"], x["Grid Size"], x["Block Size"], float(x["Metric Value"])) for x in csv.DictReader(lines)]
    idx = [i for i, x in enumerate(rows) if marker in x[0]]
    if len(idx) < 2:
        a, b = 0, len(rows)
    else:
        a, b = idx[which] + 1, idx[which + 1] + 1
    step = rows[a:b]
    tot = sum(x[3] for x in step)
    agg = collections.OrderedDict()
    for k, g, bl, t in step:
        k = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", k).split("(")[0][:80]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += t
    print(f"# ncu launch list summary: {path}\n")
    print(f"launches {a}..{b - 1} of {len(rows)} = one train step: {len(step)} launches, {tot / 1e3:.1f} us summed device time "
          "(cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) as a markdown table of the counters the roofline discussion uses: duration, DRAM bytes
and throughput, tensor-pipe activity, shared-memory wavefronts (LSU vs tensor core), occupancy, top warp-stall reasons.

    python tools/ncu_full_summary.py gpurun_out/x.ncu-rep > profiles/rXX_x_ncu_full.md
"""
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (realtime)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor hmma sub-pipe active %"),
    ("sm__inst_executed_pipe_uniform_realtime.avg.pct_of_peak_sustained_elapsed", "uniform pipe (UTCHMMA/UTMALDG issue) %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts, LSU % of peak"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts, tensor core % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts (LSU)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if re.match(r"smsp__average_warps?_issue_stalled_.*_per_issue_active", h) or re.match(r"smsp__average_warp_latency_issue_stalled_.*\.pct", h)]
    if not stall:
        stall = [h for h in hdr if "issue_stalled" in h and h.endswith(".pct")]
    print(f"# `ncu --set full --clock-control none` summary of {rep.split('/')[-1]} (per launch; replayed ~40x, cold cache: compare ratios, not times)\n")
    for r in data:
        name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", r[col["Kernel Name"]]).split("(")[0].replace("void ", "")
        print(f"## `{name}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n")
        print("| counter | value |\n|---|---:|")
        for key, label in WANT:
            if key in col and r[col[key]] != "":
                print(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |")
        st = []
        for h in stall:
            try:
                st.append((float(r[col[h]].replace(",", "")), h))
            except ValueError:
                pass
        st.sort(reverse=True)
        if st:
            print("\ntop warp-stall reasons: " + "; ".join(f"{re.sub(r'smsp__average_warps?_(latency_)?issue_stalled_|_per_issue_active.*|\\.pct', '', h)} {v:.1f}" for v, h in st[:5]))
        print()


if __name__ == "__main__":
    main()

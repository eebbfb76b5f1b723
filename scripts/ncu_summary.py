#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

    python scripts/ncu_summary.py launches gpurun_out/r3_launches.csv  > profiles/r1_launches_<tag>.txt
    python scripts/ncu_summary.py full     gpurun_out/r3_prof.ncu-rep  > profiles/r1_full_<tag>.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    unit = ""
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        unit = r[ui]
        name = r[ki].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES)")
    print(f"# source: {path}; {sum(cnt.values())} launches, total {s / 1e6:.3f} ms")
    print(f"{'kernel':62s} {'launches':>8s} {'total_' + unit:>16s} {'avg_' + unit:>14s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k[:62]:62s} {cnt[k]:8d} {v:16.0f} {v / cnt[k]:14.0f} {v / s * 100:6.2f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on; source: {path}")
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:88s} {r[i]:>18s} {units[i]}")
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        print(f"  traffic = dram read + write = {r[rd]} {units[rd]} + {r[wr]} {units[wr]}")
        print()


def traffic(path, summary_file="", queries="10000000"):
    """Write profiles/ncu_dominant_kernel.json (read by bench.py): dram bytes and warp instructions of one launch of the
    LAST kernel in the capture, with the hash of the kernel sources and the commit it was taken on."""
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, r = rows[0], rows[-1]
    val = lambda m: float(r[hdr.index(m)].replace(",", ""))
    units = rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = val("dram__bytes_read.sum") * scale[units[hdr.index("dram__bytes_read.sum")]]
    wr = val("dram__bytes_write.sum") * scale[units[hdr.index("dram__bytes_write.sum")]]
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    d = {"kernel": r[hdr.index("Kernel Name")], "queries": int(queries), "dram_bytes_read": rd, "dram_bytes_write": wr,
         "dram_bytes": rd + wr, "inst_executed": val("smsp__inst_executed.sum"), "duration_ns_under_ncu": val("gpu__time_duration.sum"),
         "source_sha1": bench.kernel_source_hash(), "commit": commit, "summary_file": summary_file}
    json.dump(d, open(bench.NCU_FILE, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])

#!/usr/bin/env python
"""Turns the ncu outputs of one GPU session into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py launches gpurun_out/launches_X.csv   > profiles/rN_launches_cfg2.md
  python scripts/summarize_profiles.py ncu      gpurun_out/prof_X.ncu-rep    > profiles/rN_ncu_cfg2.md
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        k = d["Kernel Name"].split("(")[0][:70]
        if "pg::" not in k:
            continue   # torch kernels of the synthetic data generator (bench tooling), not the library
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |")


def ncu(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    seen = collections.Counter()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0]
        seen[name] += 1
        if seen[name] > 2:
            continue
        print(f"\n## `{name}` (launch {seen[name]})\n\n| metric | value | unit |\n|---|---:|---|")
        for m, label in METRICS:
            if m in d and d[m] != "":
                print(f"| {label} (`{m}`) | {d[m]} | {unit.get(m, '')} |")


if __name__ == "__main__":
    {"launches": launches, "ncu": ncu}[sys.argv[1]](sys.argv[2])

#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py launches <launches.csv>          aggregated launch list (share of step)
  python tools/ncu_summary.py rep <file.ncu-rep> [regex]       key metrics per captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
    "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if "Metric Value" not in row or not row.get("Kernel Name"):
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        agg.setdefault(k, []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# source: {path}  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':70s} {'n':>4s} {'mean_us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[:70]:70s} {len(v):4d} {sum(v) / len(v) / 1e3:10.2f} {sum(v) / tot:7.3f}")


def rep(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# source: {path}  (ncu --set full --clock-control none --import-source on)")
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        if pat and not re.search(pat, d["Kernel Name"]):
            continue
        print(f"\n## {d['Kernel Name'][:110]}")
        for k in KEYS:
            if k in d:
                print(f"  {k:90s} {d[k]:>18s} {units[hdr.index(k)]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        rep(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)

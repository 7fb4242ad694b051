#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/: launch list CSV -> per-kernel table; .ncu-rep -> key metrics."""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "lts__t_requests_srcunit_tex_op_red.sum",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg.setdefault(r["Kernel Name"][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("%-72s %5s %11s %12s %6s" % ("kernel", "n", "mean_us", "total_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-72s %5d %11.1f %12.1f %5.1f%%" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(out.splitlines()))
    hdr, units = rd[0], rd[1]
    for row in rd[2:]:
        print("== %s" % row[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-70s %s %s" % (w, row[i], units[i]))


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else rep)(sys.argv[2])

#!/usr/bin/env python
"""Summaries of ncu artefacts for profiles/:
    ncu_summarize.py list  launches.csv          per-kernel count / mean duration / share (and DRAM bytes when captured)
    ncu_summarize.py full  capture.ncu-rep       selected metrics of every captured launch (reads it with `ncu -i ... --page raw --csv`)
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def rows_of(text):
    lines = text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))


def short(name, n=60):
    return (name[:n - 1] + "~") if len(name) > n else name


def do_list(path):
    rows = rows_of(open(path).read())
    agg = OrderedDict()
    for r in rows:
        k = r["Kernel Name"]
        a = agg.setdefault(k, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0})
        m, v, u = r["Metric Name"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        if m == "gpu__time_duration.sum":
            a["n"] += 1
            a["t"] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        elif m.startswith("dram__bytes"):
            a["rd" if "read" in m else "wr"] += v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    tot = sum(a["t"] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        extra = "  rd=%8.1f MB wr=%8.1f MB (per launch)" % (a["rd"] / a["n"], a["wr"] / a["n"]) if a["rd"] + a["wr"] > 0 else ""
        print("%-60s n=%4d avg=%9.1f us share=%5.1f%%%s" % (short(k), a["n"], a["t"] / a["n"], 100 * a["t"] / tot, extra))


def do_full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rd = csv.reader(io.StringIO("\n".join(lines[start:])))
    head, units = next(rd), next(rd)
    for row in rd:
        d = dict(zip(head, row))
        print("----\n  Kernel Name = %s" % short(d.get("Kernel Name", "?"), 100))
        for k in KEEP:
            if k in d:
                print("  %s = %s %s" % (k, d[k], units[head.index(k)]))


if __name__ == "__main__":
    (do_list if sys.argv[1] == "list" else do_full)(sys.argv[2])

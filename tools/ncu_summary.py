#!/usr/bin/env python3
"""Summarise ncu output for profiles/: key metrics of a --set full report, or a per-kernel share table of a
`--metrics gpu__time_duration.sum` launch list (CSV).  Usage:
    tools/ncu_summary.py report gpurun_out/prof.ncu-rep > profiles/rN_kernel.txt
    tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rN_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second"]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name}")
        for i, h in enumerate(hdr):
            if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                print(f"{h:95s} {vals[i]:>18s} {units[i]}")


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = defaultdict(float); cnt = defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        k = r["Kernel Name"].split("(")[0]
        tot[k] += ms; cnt[k] += 1
    total = sum(tot.values())
    print(f"{'kernel':70s} {'launches':>8s} {'total ms':>12s} {'avg ms':>10s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k[:70]:70s} {cnt[k]:8d} {v:12.4f} {v / cnt[k]:10.4f} {100 * v / total:6.1f}%")


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2])

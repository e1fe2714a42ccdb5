#!/usr/bin/env python3
"""Summarise ncu captures for profiles/:
    python tools/ncu_summary.py launches <launches.csv>            -> per-kernel-class totals / shares (markdown)
    python tools/ncu_summary.py full <a.ncu-rep> [b.ncu-rep ...]   -> key counters per captured launch (markdown)
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]


def short(name):
    m = re.search(r"(cheb_blocked_kernel|stencil_tma_pre_kernel|stencil_tma_kernel|stencil_kernel|pointwise_kernel)<([^>]*(?:<[^>]*>)?[^>]*)>", name)
    if m:
        return m.group(1) + "<" + m.group(2) + ">"
    return name.split("(")[0][:60]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f} % |")


def full(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        h, u = rows[0], rows[1]
        for r in rows[2:]:
            print(f"\n### `{short(r[h.index('Kernel Name')])}`  ({p.split('/')[-1]})\n\n| counter | value | unit |\n|---|---:|---|")
            for k in KEYS:
                if k in h:
                    print(f"| {k} | {r[h.index(k)]} | {u[h.index(k)]} |")


def traffic(cells, pairs):
    """traffic <cells> key=report.ncu-rep ... -> JSON for profiles/rNN_traffic.json (what bench.py's roofline.traffic reads)"""
    import json
    out = {"_comment": "DRAM traffic per launch from `ncu --set full --clock-control none` captures (one launch each); bench.py scales "
                       "it by cells per launch for roofline.traffic"}
    for pr in pairs:
        key, path = pr.split("=", 1)
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        h, u, r = rows[0], rows[1], rows[2]

        def val(name):
            v, unit = float(r[h.index(name)].replace(",", "")), u[h.index(name)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        out[key] = {"kernel": short(r[h.index("Kernel Name")]), "cells": cells, "dram_bytes_read": val("dram__bytes_read.sum"),
                    "dram_bytes_write": val("dram__bytes_write.sum"), "report": path.split("/")[-1]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(int(sys.argv[2]), sys.argv[3:])
    else:
        full(sys.argv[2:])

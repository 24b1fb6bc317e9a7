#!/usr/bin/env python
"""Turn `ncu --set full` captures into profiles/traffic.json, the record bench.py reads `roofline.traffic` from.

    python tools/ncu_traffic.py gpurun_out/detect_f32.ncu-rep [more.ncu-rep ...] [--round r2]

For every distinct kernel in the reports: per-launch dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration
(median over the captured launches), registers, and the sha256 of csrc/peaks.cu at capture time (bench.py reports whether
the capture matches the source it is running).  Kernel names are normalised to `name<template args>`.  Next to it, for
every report, profiles/<report>_ncu_raw.txt: the selected metrics of each captured launch (`--raw-only` skips traffic.json).
"""
import csv
import hashlib
import json
import os
import re
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    x = float(v.replace(",", ""))
    return x * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1)


def short_name(full):
    m = re.match(r"(?:void\s+)?(?:snb::)?([A-Za-z_0-9]+)(<.*>)?\(", full)
    if not m:
        return full
    return m.group(1) + (m.group(2) or "").replace(" ", "").replace("(int)", "")


RAW_METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
               "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
               "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
               "launch__block_size", "launch__cluster_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
               "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
               "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
               "smsp__inst_executed.sum", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
               "smsp__cycles_active.avg", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def write_raw(rep, hdr, units, rows):
    col = {h: i for i, h in enumerate(hdr)}
    out = os.path.join(ROOT, "profiles", os.path.basename(rep).replace(".ncu-rep", "_ncu_raw.txt"))
    with open(out, "w") as f:
        for k, r in enumerate(rows):
            if k:
                f.write("---\n")
            f.write(f"{'Kernel Name':<75} {r[col['Kernel Name']][:70]} \n")
            for m in RAW_METRICS:
                if m in col:
                    f.write(f"{m:<75} {r[col[m]]} {units[col[m]]}\n")
    return out


def main():
    raw_only = "--raw-only" in sys.argv
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rnd = "r2"
    if "--round" in sys.argv:
        rnd = sys.argv[sys.argv.index("--round") + 1]
        args = [a for a in args if a != rnd]
    with open(os.path.join(ROOT, "sleap_nn_b200", "csrc", "peaks.cu"), "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()[:16]
    path = os.path.join(ROOT, "profiles", "traffic.json")
    rec = json.load(open(path)) if os.path.isfile(path) else {}
    for rep in args:
        hdr, units, rows = rows_of(rep)
        col = {h: i for i, h in enumerate(hdr)}
        print("wrote", write_raw(rep, hdr, units, rows))
        if raw_only:
            continue
        by = {}
        for r in rows:
            by.setdefault(short_name(r[col["Kernel Name"]] if "Kernel Name" in col else r[4]), []).append(r)
        for name, rs in by.items():
            rd = [to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) for r in rs]
            wr = [to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]) for r in rs]
            us = [to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]]) for r in rs]
            rec[name] = {"dram_bytes_read": statistics.median(rd), "dram_bytes_write": statistics.median(wr),
                         "ncu_time_us": statistics.median(us), "launches_captured": len(rs),
                         "registers": int(rs[0][col["launch__registers_per_thread"]]),
                         "grid": rs[0][col["launch__grid_size"]], "block": rs[0][col["launch__block_size"]],
                         "source_sha16": sha,
                         "capture": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "_ncu_raw.txt") +
                                    " (selected metrics of an ncu --set full --clock-control none capture)"}
            print(name, rec[name])
    if not raw_only:
        json.dump(rec, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

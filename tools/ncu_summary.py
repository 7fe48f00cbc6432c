#!/usr/bin/env python
"""Summarise ncu outputs (read here, no GPU needed) into small text files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv          -> per-kernel totals / shares
  python tools/ncu_summary.py full gpurun_out/prof_fast_r1.ncu-rep         -> key metrics of a --set full capture
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg"]


def launches(path, lo=None, hi=None, label=None):
    """per-kernel totals of an ncu launch list; lo/hi restrict to a range of the library's own launches (hyorb:: kernels in
    launch order, setup kernels k_pattern_to_float / k_repack excluded)"""
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    recs = [(r[ki].split("(")[0], float(r[vi].replace(",", ""))) for r in rows[start:] if len(r) > vi]
    if lo is not None:
        own = [x for x in recs if "hyorb::" in x[0] and "k_pattern" not in x[0] and "k_repack" not in x[0]]
        recs = own[lo:hi]
    for k, v in recs:
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}{' ' + label if label else ''}: gpu__time_duration.sum per kernel (ns), cold-cache serialised replay -- compare SHARES")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:70]:70s} launches={len(v):4d} total_us={sum(v) / 1e3:10.1f} mean_us={sum(v) / len(v) / 1e3:9.1f} share={sum(v) / tot:.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print(f"# {path}: {vals[hdr.index('Kernel Name')][:100]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:75s} {units[i]:16s} {vals[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches" and len(sys.argv) > 4:
        launches(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5] if len(sys.argv) > 5 else None)
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

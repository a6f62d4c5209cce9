#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report: one row per kernel (first launch of each name), then the hottest
source lines of the chosen kernels.  usage: tools/ncu_summary.py <report.ncu-rep> [kernel ...] > profiles/xxx.md"""
import csv, subprocess, sys
rep = sys.argv[1]
hot = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
cols = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
        ("smsp__inst_executed.sum", "warp instr")]
idx = [(h.index(c), n) for c, n in cols if c in h]
print("# ncu --set full --clock-control none: %s\n" % rep.split("/")[-1])
print("Per-launch values of the first captured launch of each kernel (cold caches, serialised by the profiler: compare shares,")
print("not absolutes; bench.py's CUDA-event numbers are the timing of record).\n")
print("| " + " | ".join(n + (" [%s]" % units[i] if units[i] and n not in ("kernel",) else "") for i, n in idx) + " |")
print("|" + "---|" * len(idx))
seen = {}
for r in rows[2:]:
    name = r[idx[0][0]].split("(")[0]
    key = name + r[idx[1][0]]
    if name in seen and seen[name] >= (7 if "pyr" in name else (2 if "select" in name else 1)):
        continue
    seen[name] = seen.get(name, 0) + 1
    vals = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = name
        else:
            try:
                f = float(v.replace(",", ""))
                v = ("%.0f" % f) if f >= 1000 or f == int(f) else ("%.3g" % f)
            except ValueError:
                pass
        vals.append(v)
    print("| " + " | ".join(vals) + " |")
for k in hot:
    out = subprocess.run([sys.executable, __file__.replace("ncu_summary", "ncu_lines"), rep, k, "14"], capture_output=True, text=True).stdout
    print("\n## hottest source lines: %s\n\n```\n%s```" % (k, out))

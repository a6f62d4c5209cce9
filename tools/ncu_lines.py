#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py <report> <kernel-name> [top-n] [launch-skip]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
by = 1 if (len(sys.argv) > 5 and sys.argv[5] == "ins") else 0
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, data = None, None, []
for r in rows:
    if len(r) >= 2 and r[0].startswith("File"):
        fname = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; sa = r.index("# Samples"); ie = r.index("Instructions Executed"); continue
    if hdr and len(r) > ie and r[0].strip().isdigit():
        try:
            data.append((int(r[sa]), int(r[ie]), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
ts, ti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
print("kernel %s: %d stall samples, %d warp instructions over %d source lines" % (kern, ts, ti, len(data)))
for d in sorted(data, key=lambda x: -x[by])[:top]:
    print("%5.1f%% smp %5.1f%% ins  %-22s:%-4d %s" % (100 * d[0] / ts, 100 * d[1] / ti, d[2], d[3], d[4]))

#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, total, share, average.
usage: tools/launch_summary.py launches.csv"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0].replace("dvo::", "").replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.3f | %.1f |" % (k, a[0], a[1], a[1] / tot, a[1] / a[0]))
print("\ntotal %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.
usage: python tools/summarize_launches.py gpurun_out/x_launches.csv [skip_launches] > profiles/x_launches_summary.md"""
import csv, sys, collections, re

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).strip()
    name = name.split("::")[-1]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1000.0 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1000.0)
    rows.append((name, us, row.get("Grid Size", ""), row.get("Block Size", "")))
rows = rows[skip:]
tot = sum(r[1] for r in rows)
agg = collections.OrderedDict()
for n, us, g, b in rows:
    a = agg.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1; a[1] += us; a[2] = max(a[2], us)
print(f"# ncu launch list summary: {path} ({len(rows)} launches after skipping {skip}, {tot/1000:.2f} ms total, cold-cache serialised)\n")
print("| kernel | launches | total us | share | avg us | max us |")
print("|---|---|---|---|---|---|")
for n, (c, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {c} | {t:.1f} | {100*t/tot:.1f}% | {t/c:.2f} | {m:.1f} |")

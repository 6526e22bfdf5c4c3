#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total ms, share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    m = re.search(r"mce_kernel_entry<mce::(?:Engine<[^>]*>::)?(\w+)", name)
    short = m.group(1) if m else re.sub(r"\(.*", "", name).split("::")[-1][:48]
    rows.append((short, float(r["Metric Value"].replace(",", "")) / 1e6, r["Grid Size"], r["Block Size"]))
tot = sum(t for _, t, _, _ in rows)
agg = defaultdict(lambda: [0, 0.0, 0.0])
for n, t, g, b in rows:
    a = agg[n]
    a[0] += 1; a[1] += t; a[2] = max(a[2], t)
print("| kernel | launches | total ms | share | max single launch ms |")
print("|---|---|---|---|---|")
for n, (c, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.3f | %.1f%% | %.3f |" % (n, c, t, 100 * t / tot, mx))
print("| **all** | %d | %.3f | 100%% | |" % (len(rows), tot))

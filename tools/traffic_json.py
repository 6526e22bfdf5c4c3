#!/usr/bin/env python
"""Turns an ncu csv with dram__bytes_read.sum / dram__bytes_write.sum per KGTable launch into profiles/traffic_*.json."""
import csv
import json
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
per = {}
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(u, 1)
    per.setdefault(r["ID"], {})[r["Metric Name"]] = v * mult
n = len(per)
rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values())
wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values())
ns = sum(p.get("gpu__time_duration.sum", 0) for p in per.values())
print(json.dumps({"kernel": "KGTable2", "launches": n, "gtable_dram_bytes_per_launch": (rd + wr) / max(n, 1), "dram_read_bytes_total": rd,
                  "dram_write_bytes_total": wr, "gpu_time_ms_total_under_ncu": ns / 1e6,
                  "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every KGTable launch of one cold pass of tests/golden/leo7.mces"}))

#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X cmd`):
per-kernel launches, total time and share. Usage: python tools/launch_list_summary.py X.csv "<command that was run>" """
import collections
import csv
import sys

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr, data = rows[0], rows[1:]
ci = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in data:
    if r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
    k = r[ci["Kernel Name"]]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"# ncu launch list summary: {cmd}")
print("# cold-cache, serialised per-launch times: compare SHARES, not absolutes")
print(f"# total launches {sum(a[0] for a in agg.values())}, total kernel time {tot:.3f} ms")
print("kernel,launches,total_ms,share,avg_us")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{k}\",{a[0]},{a[1]:.3f},{a[1] / tot:.4f},{1e3 * a[1] / a[0]:.1f}")

#!/usr/bin/env python
"""Aggregate an ncu launch list (gpu__time_duration.sum per launch) by kernel name."""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    unit = r[ui]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("owc::", "")
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values()) or 1.0
print(f"{'kernel':58s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} {n:8d} {us:10.1f} {us/n:8.2f} {100*us/tot:5.1f}%")
print(f"{'TOTAL':58s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}")

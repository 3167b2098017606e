#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_list_summary.py <launches.csv> ["header comment"]"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Value" in r)
h = rows[hi]
kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[kn]).replace("void ", "").replace("cmda::", "").strip()
    v = float(r[mv].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "nsecond": v / 1e3, "usecond": v, "msecond": v * 1e3}.get(r[mu], v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for c in sys.argv[2:]:
    print("# " + c)
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {n:8d} {us:12.1f} {us / n:10.1f} {100 * us / tot:6.1f}%")

#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total,
share (cold-cache, serialised: compare SHARES with bench.py's breakdown, not absolutes)."""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    unit = r[hdr.index("Metric Unit")]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
    n = r[ik].replace("<unnamed>::", "").replace("void ", "")
    n = n[:n.index("(")] if "(" in n else n
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v for _, v in agg.values())
print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n[:70]}` | {c} | {v:.1f} | {100 * v / tot:.1f}% |")
print(f"| **all** | {sum(c for c, _ in agg.values())} | {tot:.1f} | 100% |")

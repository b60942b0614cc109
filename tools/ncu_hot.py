#!/usr/bin/env python
"""Hottest SASS lines (stall samples) of one kernel in an .ncu-rep, with the dominant stall reason:
    python tools/ncu_hot.py rep.ncu-rep kernel-regex [N]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name",')
rows = list(csv.reader(blocks[1].splitlines()))
hdr = rows[1]
si = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) <= si:
        continue
    try:
        s = int(r[si])
    except ValueError:
        continue
    top = max(stall_cols, key=lambda i: int(r[i] or 0))
    data.append((s, idx, r[1].strip(), hdr[top], int(r[top] or 0)))
tot = sum(d[0] for d in data)
print(f"total samples {tot}, instructions {len(data)}")
for s, idx, src, reason, cnt in sorted(data, reverse=True)[:n]:
    print(f"{100 * s / tot:5.1f}%  #{idx:5d}  {src[:70]:70s} {reason}={cnt}")

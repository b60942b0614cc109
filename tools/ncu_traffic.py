#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged per kernel name, from one or more .ncu-rep files
(`ncu --set full`) -> profiles/ncu_traffic.json, the `roofline.traffic` source of bench.py:
    python tools/ncu_traffic.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/ncu_traffic.json"""
import csv, json, subprocess, sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = defaultdict(lambda: [0, 0.0])
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    for r in rows[2:]:
        name = r[ik].replace("<unnamed>::", "").replace("void ", "")
        name = name[:name.index("(")] if "(" in name else name
        try:
            b = float(r[ir].replace(",", "")) * UNIT.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * UNIT.get(units[iw], 1.0)
        except ValueError:
            continue
        acc[name][0] += 1
        acc[name][1] += b
print(json.dumps({k: round(v[1] / v[0]) for k, v in sorted(acc.items())}, indent=1))

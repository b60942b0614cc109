#!/usr/bin/env python
"""Instruction-count table per kernel of cnn_b200/libcnn_b200.so (cuobjdump -sass): what proves the Blackwell-native
paths (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UBLKCP = TMA tensor / bulk copies, FFMA2 = packed fp32 FMA).
    python tools/sass_table.py > profiles/r02_sass_table.md"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cnn_b200", "libcnn_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTMAPF", "FFMA2", "FFMA", "HMMA", "LDS", "STS", "LDG", "STG", "SYNCS", "BAR", "SHFL", "ATOM"]
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    mangled = f.split("\n")[0].strip()
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = name[:name.index("(")] if "(" in name else name
    c = collections.Counter()
    total = 0
    for line in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            total += 1
            op = m.group(2)
            for k in KEYS:
                if op == k or (k in ("LDS", "STS", "LDG", "STG", "ATOM", "BAR", "SHFL", "SYNCS") and op.startswith(k)):
                    c[k] += 1
                    break
    rows.append((name, total, c))
print(f"# SASS instruction counts per kernel of `{os.path.relpath(so, ROOT)}` (static counts, `cuobjdump -sass`)\n")
print("| kernel | instr | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for name, total, c in sorted(rows):
    print(f"| `{name[:80]}` | {total} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
tot = collections.Counter()
for _, _, c in rows:
    tot.update(c)
print("| **all kernels** | " + str(sum(r[1] for r in rows)) + " | " + " | ".join(str(tot[k]) for k in KEYS) + " |")

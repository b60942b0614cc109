#!/usr/bin/env python
"""Warp-stall breakdown (pc sampling) + pipe utilisation per kernel of an .ncu-rep:
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [kernel-substring]"""
import csv, subprocess, sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = sys.argv[2] if len(sys.argv) > 2 else ""
seen = set()
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")[:50]
    if want not in name or name in seen:
        continue
    seen.add(name)
    st = []
    for k, v in d.items():
        if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
            try:
                st.append((k[len("smsp__pcsamp_warps_issue_stalled_"):], float(v.replace(",", ""))))
            except ValueError:
                pass
    tot = sum(v for _, v in st) or 1.0
    print(f"{name}: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(st, key=lambda kv: -kv[1])[:7]))
    for k, lbl in [("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
                   ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe %"),
                   ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu inst %"),
                   ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
                   ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
                   ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps/cycle"),
                   ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
                   ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
                   ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
                   ("lts__t_sector_hit_rate.pct", "L2 hit %")]:
        print(f"    {lbl}: {d.get(k)}")

#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` into the per-kernel table committed under profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.md"""
import csv, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__inst_executed.sum", "warp inst")]

def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{rep}` (--set full, --clock-control none)\n")
    print("| kernel | " + " | ".join(k for _, k in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "?").replace("<unnamed>::", "").replace("void ", "")[:60]
        cells = []
        for key, _ in KEYS:
            v = d.get(key, "")
            u = units[hdr.index(key)] if key in hdr else ""
            try:
                cells.append(f"{float(v.replace(',', '')):.4g} {u}".strip())
            except ValueError:
                cells.append(v or "-")
        print(f"| `{name}` | " + " | ".join(cells) + " |")

if __name__ == "__main__":
    main()

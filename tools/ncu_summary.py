#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per kernel launch with duration, DRAM bytes,
DRAM %, achieved occupancy, registers and the top warp-stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
if not stall:
    stall = [h for h in hdr if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
units = rows[1]
def g(r, k, d="-"):
    return r[col[k]] if k in col else d
_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
def mb(r, k):
    """a byte metric in MB whatever unit ncu chose for the column (it scales per report)"""
    return float(g(r, k, "0") or 0) * _MB.get(units[col[k]], 1.0) if k in col else 0.0
_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
def us(r, k):
    return float(g(r, k, "0") or 0) * _US.get(units[col[k]], 1.0)
print(f"{'kernel':44s} {'grid':>8s} {'us':>8s} {'rdMB':>8s} {'wrMB':>8s} {'dram%':>6s} {'occ%':>6s} {'regs':>4s} {'ipc':>5s}  top stalls")
for r in rows[2:]:
    name = g(r, "Kernel Name")[:44]
    st = sorted(((float(r[col[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:3]
    grid = g(r, "Grid Size").strip("()").split(",")[0]
    print(f"{name:44s} {grid:>8s} {us(r,'gpu__time_duration.sum'):8.1f} {mb(r,'dram__bytes_read.sum'):8.1f} {mb(r,'dram__bytes_write.sum'):8.1f} "
          f"{float(g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} {float(g(r,'sm__warps_active.avg.pct_of_peak_sustained_active')):6.1f} {g(r,'launch__registers_per_thread'):>4s} "
          f"{float(g(r,'sm__inst_executed.avg.per_cycle_elapsed', '0') or 0):5.2f}  " + ", ".join(f"{n}={v:.1f}" for v, n in st))

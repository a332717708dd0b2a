#!/usr/bin/env python
"""Per-kernel SASS histogram from an `ncu --page source --csv` export: executed warp instructions by opcode,
stall samples by opcode, shared-memory wavefronts (excess = bank conflicts).  usage: ncu_src_hist.py src.csv [kernel regex]"""
import collections
import csv
import re
import sys

pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
kernel, hdr, col = None, None, {}
stats = collections.OrderedDict()
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "Kernel Name":
        kernel, hdr = r[1], None
        continue
    if r[0] == "Address":
        hdr = r
        col = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or kernel is None or (pat and not pat.search(kernel)):
        continue
    src = r[col["Source"]].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    op = op.rstrip(";")
    key = ".".join(op.split(".")[:2])
    d = stats.setdefault(kernel, dict(inst=collections.Counter(), stall=collections.Counter(), wave=0, ideal=0, total=0, samples=0))
    n = int(r[col["Instructions Executed"]] or 0)
    s = int(r[col["# Samples"]] or 0)
    d["inst"][key] += n
    d["stall"][key] += s
    d["total"] += n
    d["samples"] += s
    d["wave"] += int(r[col["L1 Wavefronts Shared"]] or 0)
    d["ideal"] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
for k, d in stats.items():
    print(f"== {k[:110]}\n   warp instructions {d['total']}, stall samples {d['samples']}, smem wavefronts {d['wave']} (ideal {d['ideal']})")
    print("   by instructions: " + ", ".join(f"{o}={n} ({100*n/max(1,d['total']):.0f}%)" for o, n in d["inst"].most_common(14)))
    print("   by stall samples: " + ", ".join(f"{o}={n} ({100*n/max(1,d['samples']):.0f}%)" for o, n in d["stall"].most_common(10)))

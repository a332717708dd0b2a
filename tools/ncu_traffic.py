#!/usr/bin/env python
"""profiles/kernel_traffic.json from an `ncu --page raw --csv` export: DRAM bytes (read + write) and duration of one
launch per kernel, keyed by the names bench.py looks up.  usage: ncu_traffic.py raw.csv out.json "source tag" """
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
KEYS = {"student_step_kernel<0>": "student_step_adam", "student_step_kernel<(int)0>": "student_step_adam",
        "grad_check_kernel": "grad_check", "ema_multi_kernel<float": "ema_multi", "adain_warp_kernel<float": "adain_warp_f32",
        "dp_gather_ema_kernel": "dp_gather_ema", "dp_reduce_step_kernel": "dp_reduce_step", "loss_step_kernel": "loss_step",
        "decode_tma_kernel<float": "decode_tma_f32", "pck_tma_kernel": "pck_tma", "gauss_target_multi_kernel": "gauss_target_multi"}
out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    for pat, key in KEYS.items():
        if pat in name:
            def val(metric, table):
                return float(r[col[metric]] or 0) * table.get(units[col[metric]], 1.0)
            rec = dict(kernel=name[:120], dram_bytes_per_launch=val("dram__bytes_read.sum", _B) + val("dram__bytes_write.sum", _B),
                       dram_read=val("dram__bytes_read.sum", _B), dram_write=val("dram__bytes_write.sum", _B),
                       duration_us=val("gpu__time_duration.sum", _US), grid=r[col["Grid Size"]], source=sys.argv[3])
            if key not in out or rec["dram_bytes_per_launch"] > out[key]["dram_bytes_per_launch"]:
                out[key] = rec      # the largest launch of a family (the microbench size)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: (round(v["dram_bytes_per_launch"] / 1e6, 1), round(v["duration_us"], 1)) for k, v in out.items()}))

#!/bin/bash
# teacher chain in one launch (udape_rewarp_decode_select): tests, microbench rows, bench line
TAG=${1:-r02ba}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/microbench.py --only rewarp,decode --configs C2,C4,C5 --out $O/${TAG}_mb.json 2>&1 | grep -E "rewarp_fwd f32|rewarp f32|kth-select" | cut -c1-110
timeout 600 python bench.py --skip-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["step_algorithmic_bytes"], d["roofline"]["step_frac_of_peak"], d["variants"]["ema"])
print(d["roofline"]["step_bytes_by_kernel"])
PY

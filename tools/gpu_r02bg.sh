#!/bin/bash
# stream priorities of the step with the Adam tail: which of {AdaIN, tail} should the block scheduler drain first?
O=gpurun_out; mkdir -p $O
for combo in "-1 0" "0 -1" "-1 -1" "0 0" "-1 -2" "-1 0"; do
  set -- $combo
  UDAPE_ADAIN_PRIORITY=$1 UDAPE_TAIL_PRIORITY=$2 timeout 300 python bench.py --skip-cpu-baseline --steps 100 --warmup 10 > $O/r02bg_bench.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("$O/r02bg_bench.json").read().strip().splitlines()[-1])
print("adain $1 tail $2:", round(d["value"]), d["ms_per_step"], "ema-only", d["variants"]["ema"]["ms_per_step"])
PY
done

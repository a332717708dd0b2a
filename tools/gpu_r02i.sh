#!/bin/bash
TAG=${1:-r02i}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_rewarp.py tests/test_gpu_hotpath.py tests/test_gpu_golden.py -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
timeout 600 python tools/microbench.py --only rewarp --configs C2,C5 --out $O/${TAG}_microbench.json > $O/${TAG}_microbench.log 2>&1; echo "microbench exit $?"
grep -v "^wrote" $O/${TAG}_microbench.log | tail -14
echo "--- 256-thread kernels (UDAPE_REWARP_WIDE=0)"
UDAPE_REWARP_WIDE=0 timeout 600 python tools/microbench.py --only rewarp --configs C5 --out $O/${TAG}_microbench_w0.json 2>&1 | grep -v "^wrote" | tail -7
timeout 600 python bench.py --skip-cpu-baseline --no-variants > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read())
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac_of_peak"], d["gpu_launches"])
PY

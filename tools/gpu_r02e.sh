#!/bin/bash
# r02e: GPU suite, microbench groups (rewarp, adain, target) incl. AdaIN warps-per-CTA A/B, bench N=1
TAG=${1:-r02e}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -12 $O/${TAG}_pytest.log
timeout 600 python tools/microbench.py --only rewarp,adain,target --configs C2,C5 --adain-n 32,64 --out $O/${TAG}_microbench.json > $O/${TAG}_microbench.log 2>&1; echo "microbench exit $?"
grep -v "^wrote" $O/${TAG}_microbench.log | tail -60
echo "--- adain 4 warps per CTA"
UDAPE_ADAIN_WARPS=4 timeout 300 python tools/microbench.py --only adain --adain-n 32,64 --configs "" --out $O/${TAG}_mb_adain_w4.json 2>&1 | grep -v "^wrote" | tail -12
echo "--- rewarp bwd CTAs per SM 3 / 12"
UDAPE_REWARP_BWD_CTAS=3 timeout 300 python tools/microbench.py --only rewarp --configs C5 --out $O/${TAG}_mb_rw3.json 2>&1 | grep "bwd" | tail -3
UDAPE_REWARP_BWD_CTAS=12 timeout 300 python tools/microbench.py --only rewarp --configs C5 --out $O/${TAG}_mb_rw12.json 2>&1 | grep "bwd" | tail -3
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read())
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["tail_kernels_ms"], d["roofline"]["step_frac_of_peak"])
print(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"], d["eager_cuda_baseline"])
print(d["variants"])
PY
tail -3 $O/${TAG}_bench.err

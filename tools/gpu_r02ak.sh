#!/bin/bash
# push backward: cost-balanced pass order (UDAPE_RW_PASS_FIXED = fixed cost of a pass in planes)
TAG=${1:-r02ak}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_rewarp.py tests/test_gpu_hotpath.py -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
for n in 0 1 2 4 1000; do
  echo "== UDAPE_RW_PASS_FIXED=$n"
  UDAPE_RW_PASS_FIXED=$n timeout 300 python tools/microbench.py --only rewarp --configs C4,C5 --out $O/${TAG}_mb_$n.json 2>&1 | grep "rewarp_bwd f16 (plan" | cut -c1-110
done

#!/bin/bash
# push-plan backward: parity tests, then the re-warp microbench rows for a few grid sizes
TAG=${1:-r02o}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_rewarp.py tests/test_gpu_hotpath.py -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
for v in 3 2 6 12; do
  echo "== PER_SM=$v"
  UDAPE_REWARP_BWD_PER_SM=$v timeout 600 python tools/microbench.py --only rewarp --configs C2,C5 \
     --out $O/${TAG}_mb_$v.json 2>&1 | grep -E "plan|student" 
done

#!/bin/bash
# push-plan backward: thread-count variants — parity tests and the re-warp microbench rows
TAG=${1:-r02o}
O=gpurun_out; mkdir -p $O
for v in "512 2" "256 2"; do
  set -- $v
  echo "== NT=$1 PER_SM=$2"
  UDAPE_REWARP_BWD_NT=$1 UDAPE_REWARP_BWD_PER_SM=$2 timeout 300 python -m pytest tests/test_gpu_rewarp.py -q -x 2>&1 | tail -1
  UDAPE_REWARP_BWD_NT=$1 UDAPE_REWARP_BWD_PER_SM=$2 timeout 300 python tools/microbench.py --only rewarp --configs C2,C5 --out $O/${TAG}_mb_$1_$2.json 2>&1 | grep -E "bwd f16 \(plan"
done
timeout 300 python -m pytest tests/test_gpu_hotpath.py -q -x 2>&1 | tail -1

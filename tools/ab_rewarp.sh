#!/bin/bash
# cluster-size sweep of the re-warp kernels: bash tools/ab_rewarp.sh <tag>
TAG=${1:-rw}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_rewarp.py -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
for n in ${NS:-auto 1 2 4 8}; do
  echo "== UDAPE_REWARP_CLUSTER=$n"
  if [ $n = auto ]; then unset UDAPE_REWARP_CLUSTER; else export UDAPE_REWARP_CLUSTER=$n; fi
  timeout 300 python tools/microbench.py --only rewarp --configs C2,C4,C5 --out $O/${TAG}_mb_$n.json 2>&1 | grep rewarp
done

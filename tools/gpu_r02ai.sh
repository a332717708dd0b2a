#!/bin/bash
# wide2 forward: tile shapes of a warp's gather (UDAPE_RW_TILE = 0 row walk | 2 = 8x8 px | 3 = 16x4 px | 1 | 4) x swizzle keys
TAG=${1:-r02ai}; O=gpurun_out; mkdir -p $O
for s in 1 2; do
UDAPE_RW_SWZ=$s timeout 600 python -m pytest tests/test_gpu_rewarp.py -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
done
for s in 0 1 2; do
for n in 0 2 3 4; do
  echo "== UDAPE_RW_TILE=$n UDAPE_RW_SWZ=$s"
  UDAPE_RW_SWZ=$s UDAPE_RW_TILE=$n timeout 300 python tools/microbench.py --only rewarp --configs C5 --out $O/${TAG}_mb_${n}_$s.json 2>&1 | grep "rewarp_fwd f.. (" | cut -c1-110
done
done

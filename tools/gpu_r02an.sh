#!/bin/bash
# last-CTA tails batched (loss sums, PCK flags, k-th select cached in shared memory): tests + A/B against the HEAD library
TAG=${1:-r02an}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "not rewarp and not dp" > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
for v in new head new2; do
  echo "== $v"
  if [ $v = head ]; then export UDAPE_LIB=$PWD/build/variants/head.so; else unset UDAPE_LIB; fi
  timeout 300 python tools/microbench.py --only decode,pck,loss --configs C2,C5 --out $O/${TAG}_mb_$v.json 2>&1 | grep -E "decode|mask_select|pck|fwd|loss_step" | cut -c1-100
done

#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in "UDAPE_NO_TMA=1" "UDAPE_NO_TMA=0"; do
  echo "== $v"; env $v timeout 300 python tools/microbench.py --only decode,loss,pck --configs C2,C5 --out $O/ab_$v.json | grep -v wrote
done

#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q -k adain 2>&1 | tail -2
for v in "UDAPE_NO_TMA=1" "UDAPE_PIPE_STAGES=8" "UDAPE_PIPE_STAGES=16" "UDAPE_PIPE_STAGES=24" "UDAPE_NO_TMA=1"; do
  echo "== $v"; env $v timeout 200 python tools/microbench.py --only adain --out $O/ab.json | grep -v wrote
done

#!/bin/bash
TAG=${1:-r02o}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "loss or step or hotpath or golden" 2>&1 | tail -3
timeout 300 python tools/microbench.py --only loss --configs C2,C5 --out $O/${TAG}_mb.json 2>&1 | grep -E "loss_step|cons_|mse"

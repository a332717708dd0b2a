#!/bin/bash
TAG=${1:-r02o}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/microbench.py --only rewarp --configs C2,C5 --out $O/${TAG}_microbench_rewarp.json 2>&1 | grep -E "rewarp_"
timeout 300 python bench.py --skip-cpu-baseline > $O/${TAG}_bench.json 2>$O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.json

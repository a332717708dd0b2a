#!/bin/bash
TAG=${1:-r02al}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_rewarp.py tests/test_gpu_hotpath.py -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
timeout 300 python tools/microbench.py --only rewarp --configs C2,C4,C5 --out $O/${TAG}_mb.json 2>&1 | grep "rewarp" | cut -c1-110

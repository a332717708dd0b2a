#!/bin/bash
TAG=${1:-r02m}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_rewarp.py tests/test_gpu_hotpath.py -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 600 python tools/microbench.py --only rewarp --configs C2,C5 --out $O/${TAG}_microbench.json > $O/${TAG}_microbench.log 2>&1; echo "microbench exit $?"
grep -v "^wrote" $O/${TAG}_microbench.log | tail -12

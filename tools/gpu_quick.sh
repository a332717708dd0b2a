#!/bin/bash
# quick GPU check: parity tests (bounded), then selected microbench groups
TAG=${1:-q}; GROUPS_=${2:-adain}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
timeout 300 python tools/microbench.py --only $GROUPS_ --out $O/${TAG}_microbench.json > $O/${TAG}_microbench.log 2>&1; echo "microbench exit $?"
grep -v "^wrote" $O/${TAG}_microbench.log | tail -70

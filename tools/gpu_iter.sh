#!/bin/bash
# one short GPU visit: parity tests, selected microbench groups, a bench line
# usage (under gpurun): bash tools/gpu_iter.sh <tag> <microbench groups> [pytest -k expr]
TAG=${1:-i}; GROUPS_=${2:-optim}; KEXPR=${3:-}
O=gpurun_out; mkdir -p $O
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $O/${TAG}_pytest.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1
fi
echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -25 $O/${TAG}_pytest.log
timeout 400 python tools/microbench.py --only $GROUPS_ --configs C2,C5 --out $O/${TAG}_microbench.json > $O/${TAG}_microbench.log 2>&1; echo "microbench exit $?"
grep -v "^wrote" $O/${TAG}_microbench.log | tail -40
timeout 300 python bench.py --skip-cpu-baseline --steps 200 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
python -c "import sys,json; d=json.loads(open('$O/${TAG}_bench.json').read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac_of_peak'], d['e2e']['value'])"

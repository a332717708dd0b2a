#!/bin/bash
# r02g: GPU suite, rewarp microbench (map route v2), smoke, bench C1/C4 lines
TAG=${1:-r02g}
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -12 $O/${TAG}_pytest.log
timeout 600 python tools/microbench.py --only rewarp --configs C2,C5 --out $O/${TAG}_microbench.json > $O/${TAG}_microbench.log 2>&1; echo "microbench exit $?"
grep -v "^wrote" $O/${TAG}_microbench.log | tail -30
for C in C1 C4; do
  timeout 600 python bench.py --config $C --skip-cpu-baseline --no-variants > $O/${TAG}_bench_$C.json 2> $O/${TAG}_bench_$C.err; echo "bench $C exit $?"
  python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench_$C.json").read())
print("$C", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_frac_of_peak"], d["e2e"]["value"])
PY
done

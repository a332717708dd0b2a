#!/bin/bash
TAG=${1:-s}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
for m in graph graph-serial after; do
  echo "== --ema $m"; timeout 300 python bench.py --ema $m --skip-cpu-baseline --steps 200 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac_of_peak'], d['e2e']['value'])"
done
echo "== --unfused --ema after (round-1a configuration)"; timeout 300 python bench.py --unfused --ema after --skip-cpu-baseline --steps 200 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac_of_peak'])"
echo "== no graph"; timeout 300 python bench.py --no-graph --skip-cpu-baseline --steps 200 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
timeout 600 python tools/microbench.py --out $O/${TAG}_microbench.json > $O/${TAG}_microbench.log 2>&1; grep -v wrote $O/${TAG}_microbench.log | grep -E "C5|C2|x512|Pose|clamp|f32  "

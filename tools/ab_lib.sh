#!/bin/bash
# ab_lib.sh <microbench groups> <extra microbench args...> : same-box A/B of the in-tree library against every build/variants/*.so
GROUPS_=$1; shift
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== in-tree"; timeout 300 python tools/microbench.py --only $GROUPS_ "$@" --out $O/ab_main.json | grep -v wrote
for v in build/variants/*.so; do
  echo "== $v"; UDAPE_LIB=$PWD/$v timeout 300 python tools/microbench.py --only $GROUPS_ "$@" --out $O/ab_$(basename $v .so).json | grep -v wrote
done
echo "== in-tree (again)"; timeout 300 python tools/microbench.py --only $GROUPS_ "$@" --out $O/ab_main2.json | grep -v wrote

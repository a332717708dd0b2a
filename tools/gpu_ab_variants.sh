#!/bin/bash
# same-box A/B of the in-tree library against every build/variants/*.so: gpu_ab_variants.sh <tag> <groups> <configs> <grep>
TAG=${1:-ab}; GROUPS_=${2:-loss}; CFG=${3:-C2,C5}; PAT=${4:-.}
O=gpurun_out; mkdir -p $O
echo "== in-tree"; timeout 300 python tools/microbench.py --only $GROUPS_ --configs $CFG --out $O/${TAG}_main.json 2>&1 | grep -E "$PAT" | cut -c1-100
for v in build/variants/*.so; do
  echo "== $v"; UDAPE_LIB=$PWD/$v timeout 300 python tools/microbench.py --only $GROUPS_ --configs $CFG --out $O/${TAG}_$(basename $v .so).json 2>&1 | grep -E "$PAT" | cut -c1-100
done
echo "== in-tree (again)"; timeout 300 python tools/microbench.py --only $GROUPS_ --configs $CFG --out $O/${TAG}_main2.json 2>&1 | grep -E "$PAT" | cut -c1-100

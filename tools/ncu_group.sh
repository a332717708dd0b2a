#!/bin/bash
# ncu --set full capture of one microbench group: bash tools/ncu_group.sh <tag> <group> <kernel-regex> [configs]
TAG=${1:-x}; GROUP=${2:-rewarp}; REGEX=${3:-rewarp}; CONFIGS=${4:-C2,C5}
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -o $O/${TAG}_${GROUP}_full -f \
    python tools/microbench.py --warmup 0 --iters 1 --no-flush --no-sustained --only $GROUP --configs $CONFIGS \
    --out $O/${TAG}_${GROUP}_under_ncu.json > $O/${TAG}_${GROUP}_ncu.log 2>&1
tail -3 $O/${TAG}_${GROUP}_ncu.log
ncu -i $O/${TAG}_${GROUP}_full.ncu-rep --page raw --csv > $O/${TAG}_${GROUP}_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_${GROUP}_raw.csv | tee $O/${TAG}_${GROUP}_summary.txt

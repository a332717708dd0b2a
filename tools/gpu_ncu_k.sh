#!/bin/bash
# ncu --set full of selected kernels in a microbench group; usage: gpu_ncu_k.sh TAG 'regex' group configs
TAG=$1; RE=$2; GRP=$3; CFG=${4:-C5}
O=gpurun_out; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -o $O/${TAG} -f \
    python tools/microbench.py --warmup 0 --iters 1 --no-flush --no-sustained --only $GRP --configs $CFG --adain-n 32 --out $O/${TAG}_mb.json > $O/${TAG}_ncu.log 2>&1
tail -2 $O/${TAG}_ncu.log
ncu -i $O/${TAG}.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_raw.csv | tee $O/${TAG}_summary.txt
ncu -i $O/${TAG}.ncu-rep --page source --csv > $O/${TAG}_src.csv 2>/dev/null
python tools/ncu_src_hist.py $O/${TAG}_src.csv | cut -c1-700
rm -f $O/${TAG}.ncu-rep

#!/bin/bash
# r02f: ncu --set full of the re-warp kernels (both routes) at C5 + the new kernels (student step in the bench sizes, dp world=1, targets x5)
TAG=${1:-r02f}
O=gpurun_out; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'rewarp' -o $O/${TAG}_rewarp -f \
    python tools/microbench.py --warmup 0 --iters 1 --no-flush --no-sustained --only rewarp --configs C5 --out $O/${TAG}_mb_under_ncu.json > $O/${TAG}_ncu.log 2>&1
tail -3 $O/${TAG}_ncu.log
ncu -i $O/${TAG}_rewarp.ncu-rep --page raw --csv > $O/${TAG}_rewarp_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_rewarp_raw.csv | tee $O/${TAG}_rewarp_summary.txt
ncu -i $O/${TAG}_rewarp.ncu-rep --page source --csv > $O/${TAG}_rewarp_src.csv 2>/dev/null
ls -la $O/${TAG}_* 
if [ "$(stat -c%s $O/${TAG}_rewarp.ncu-rep 2>/dev/null || echo 0)" -gt 30000000 ]; then rm -f $O/${TAG}_rewarp.ncu-rep; fi

#!/bin/bash
# ncu --set full capture of one launch of every kernel family at the microbench sizes (N=32 AdaIN, C5 heatmaps, EMA).
TAG=${1:-r01x}; shift
O=gpurun_out
mkdir -p $O
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'adain|mean_std|decode|pck|mse|cons_|loss_step|ema_multi|gauss_target|labelmap|mask_select|clamp' -o $O/${TAG}_full -f \
    python tools/microbench.py --warmup 0 --iters 1 --no-flush --no-sustained --adain-n 32 --configs C5 "$@" --out $O/${TAG}_mb_under_ncu.json > $O/${TAG}_full.log 2>&1
tail -5 $O/${TAG}_full.log
ls -la $O | tail -5

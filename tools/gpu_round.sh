#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu --set full of every kernel family.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01x}
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/${TAG}_smoke.log
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.json
python bench.py --impl reference --steps 4 --warmup 3 > $O/${TAG}_bench_reference.json 2>/dev/null; tail -c 300 $O/${TAG}_bench_reference.json
python tools/microbench.py --out $O/${TAG}_microbench.json > $O/${TAG}_microbench.log 2>&1; tail -70 $O/${TAG}_microbench.log
# launch list of the bench command (graph replay: kernels inside the graph are listed individually)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > $O/${TAG}_bench_under_ncu.log 2>&1
# full-section capture of one launch of every kernel family at the microbench sizes (NOFULL=1 skips it)
[ "${NOFULL:-0}" = "1" ] && { ls -la $O | tail -20; exit 0; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'adain|mean_std|decode|pck|mse|cons_|loss_step|ema_multi|gauss_target|labelmap|mask_select|clamp|rewarp|student_step|grad_check' -o $O/${TAG}_full -f \
    python tools/microbench.py --warmup 0 --iters 1 --no-flush --no-sustained --adain-n 32 --configs C5 --out $O/${TAG}_mb_under_ncu.json > $O/${TAG}_full.log 2>&1
# gpurun_out/ merges at most 64 MiB back: keep the raw-page CSV, drop the report itself when it is large
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_ncu_full_raw.csv 2>/dev/null
if [ "$(stat -c%s $O/${TAG}_full.ncu-rep 2>/dev/null || echo 0)" -gt 40000000 ]; then rm -f $O/${TAG}_full.ncu-rep; fi
ls -la $O | tail -20

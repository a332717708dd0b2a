#!/bin/bash
# One GPU-box visit at the end of a change set: smoke, the whole GPU suite, the bench line (+ reference arm), the full
# microbench, the eager-vs-kernel table, the ncu launch list of the bench command, an ncu --set full capture of one launch
# of every kernel family and the traffic record bench.py reads.   usage (under gpurun): bash tools/gpu_round2.sh <tag>
TAG=${1:-r02x}
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/${TAG}_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"; tail -c 400 $O/${TAG}_bench.json; echo
timeout 600 python bench.py --impl reference --steps 4 --warmup 3 > $O/${TAG}_bench_reference.json 2>/dev/null; tail -c 300 $O/${TAG}_bench_reference.json; echo
timeout 1200 python tools/microbench.py --out $O/${TAG}_microbench.json > $O/${TAG}_microbench.log 2>&1; echo "microbench exit $?"
grep -v "^wrote" $O/${TAG}_microbench.log | tail -80
timeout 900 python tools/eager_table.py --out $O/${TAG}_eager_table > $O/${TAG}_eager.log 2>&1; echo "eager table exit $?"; tail -3 $O/${TAG}_eager.log
# launch list of the bench command (graph replay: kernels inside the graph are listed individually)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --no-variants > $O/${TAG}_bench_under_ncu.log 2>&1; echo "launch list exit $?"
# one launch of every kernel family with the full section set
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'adain|mean_std|decode|pck|mse|cons_|loss_step|ema_multi|gauss_target|labelmap|mask_select|clamp|rewarp|student_step|grad_check|table_feed' \
    -o $O/${TAG}_full -f python tools/microbench.py --warmup 0 --iters 1 --no-flush --no-sustained --adain-n 32 --configs C5 \
    --out $O/${TAG}_mb_under_ncu.json > $O/${TAG}_full.log 2>&1; echo "ncu full exit $?"
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_ncu_full_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_ncu_full_raw.csv > $O/${TAG}_ncu_summary.txt 2>&1; tail -45 $O/${TAG}_ncu_summary.txt
python tools/ncu_traffic.py $O/${TAG}_ncu_full_raw.csv $O/${TAG}_kernel_traffic.json "ncu --set full, tools/gpu_round2.sh $TAG"
rm -f $O/${TAG}_full.ncu-rep
bash tools/sanitize.sh > $O/${TAG}_sanitize.log 2>&1; tail -12 $O/${TAG}_sanitize.log
ls -la $O | grep ${TAG}

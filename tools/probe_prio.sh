export PROBE_TIMELINE=1
for n in 1 2 4 8; do echo "#### plan cluster $n"; UDAPE_REWARP_PLAN_CLUSTER=$n python tools/step_probe.py 2>&1 | grep -vE "^--|ema ctas|serial|no re-warp|no AdaIN  |^no EMA" | head -24; done

export PROBE_TIMELINE=1
export UDAPE_CHAIN_PRIORITY=-2
for n in 0 1 2; do for ap in -1 0; do
echo "#### adain ctas/sm $n, adain prio $ap"; UDAPE_ADAIN_CTAS_PER_SM=$n UDAPE_ADAIN_PRIORITY=$ap python tools/step_probe.py 2>&1 | grep -E "timeline: full|adain|ema done|bwd done|join|^full step|^no EMA" | head -7
done; done
for n in 0 1 2; do echo "### adain alone ctas/sm $n"; UDAPE_ADAIN_CTAS_PER_SM=$n python tools/microbench.py --only adain --adain-n 32 --out /tmp/x.json 2>&1 | grep "adain_mix.*f32"; done

export PROBE_TIMELINE=0
for n in 1 2; do for t in 296 148 64 32; do echo "#### rewarp cluster $n bwd target $t"; UDAPE_REWARP_CLUSTER=$n UDAPE_REWARP_BWD_TARGET=$t python tools/step_probe.py 2>&1 | grep -E "^full step|^no EMA|heatmap chains alone"; done; done

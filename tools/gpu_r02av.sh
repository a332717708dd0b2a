#!/bin/bash
# fused loss step: one supervised + one consistency plane per CTA (UDAPE_LOSS_PAIR=1) against the plane-per-CTA grid
TAG=${1:-r02av}; O=gpurun_out; mkdir -p $O
UDAPE_LOSS_PAIR=1 timeout 600 python -m pytest tests -m gpu -x -q -k "loss or step or hotpath or golden" 2>&1 | tail -2
python - <<'PY'
import os, torch
import uda_poseestimation_b200 as U
from uda_poseestimation_b200 import synthetic as S
from uda_poseestimation_b200.loss import fused_losses
dev = torch.device("cuda", 0)
for (b, k, dt) in ((32, 16, torch.float16), (256, 21, torch.float16), (5, 3, torch.float32), (64, 18, torch.bfloat16)):
    ys = S.heatmaps(b, k, seed=1).to(dt).to(dev); yt = S.heatmaps(b, k, seed=2).to(dt).to(dev)
    tea = S.heatmaps(b, k, seed=3, peak=(0.3, 1.2)).to(dev)
    j, v = S.keypoints(b, k, seed=4)
    label, weight = U.generate_target_batched(j, v, (64, 64), 2, (256, 256), device=dev)
    tt = U.teacher_targets(tea, 2, 0.5, occlude_thresh=0.9, materialise=False)
    outs = []
    for flag in ("0", "1"):
        os.environ["UDAPE_LOSS_PAIR"] = flag
        losses, g1, g2 = fused_losses(ys, label, weight, yt, None, tt["tea_mask"], lambda_c=1.0, grad_scale=65536.0, tea_preds=tt["preds"], sigma=2)
        outs.append((losses.clone(), g1.clone(), g2.clone()))
    same = all(torch.equal(a, b_) for a, b_ in zip(*outs))
    print(f"pair == plane-per-CTA, bit for bit  B={b} K={k} {dt}: {same}  losses {outs[1][0].tolist()}")
PY
for f in 0 1 0 1; do
  echo "== UDAPE_LOSS_PAIR=$f"
  UDAPE_LOSS_PAIR=$f timeout 300 python tools/microbench.py --only loss --configs C2,C4,C5 --out $O/${TAG}_mb_$f.json 2>&1 | grep "loss_step (ana" | cut -c1-100
done

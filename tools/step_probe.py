"""Where does the step's time go?  Replays CUDA graphs of the hot-path step with chains left out
(profiling aid; not a bench number).  usage (GPU box): python tools/step_probe.py [--config C2]"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench as B  # noqa: E402
import uda_poseestimation_b200 as U  # noqa: E402
from uda_poseestimation_b200 import synthetic as S  # noqa: E402
from uda_poseestimation_b200.hotpath import HotPathStep, StepInputs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    cfg = S.CONFIGS[args.config]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    host = B.make_host_inputs(cfg, 1234)
    d = {n: t.to(dev) for n, t in host.items() if torch.is_tensor(t)}
    t_tea, t_stu = B.stage_tables(host, torch.float16)
    label, weight = U.generate_target_batched(d["joints"], d["vis"], (64, 64), cfg["sigma"], (256, 256), device=dev)
    alpha = torch.full((2,), 0.5, device=dev)
    shapes = S.pose_resnet_param_shapes(cfg["joints"])
    student = B.ParamBag(S.parameter_list(shapes, 7, device=dev))
    teacher = B.ParamBag([torch.empty_like(p) for p in student.parameters()])

    def timeline(title, models=None, **kw):
        inp = StepInputs(feat_src=d["feat_src"], feat_tgt_ori=d["feat_tgt_ori"], feat_tgt_tea=d["feat_tgt_tea"],
                         feat_src_ori=d["feat_src_ori"], y_s=d["y_s"], y_t_stu=d["y_t_stu"], y_t_tea=d["y_t_tea"],
                         label_s=label, weight_s=weight, alpha_s2t=alpha[0:1], alpha_t2s=alpha[1:2],
                         theta_tea=t_tea.to(dev), theta_stu=t_stu.to(dev))
        tea_m, stu_m = models if models is not None else (teacher, student)
        step = HotPathStep(tea_m, stu_m, sigma=cfg["sigma"], **kw)
        step.marks = []
        if kw.get("tail") is not None:
            kw["tail"].mark = step._mark
        step.capture(inp, include_ema=True, warmup=2)
        acc = {}
        n = 20
        for _ in range(n + 3):
            step.replay()
            torch.cuda.synchronize()
            t0 = step.marks[0][1]
            for name, ev in step.marks[1:]:
                acc.setdefault(name, []).append(t0.elapsed_time(ev) * 1e3)
        print(f"== timeline: {title} (us after start, median of {n} replays with a sync between them)")
        for name, ts in sorted(acc.items(), key=lambda kv: np.median(kv[1][3:])):
            print(f"   {name:<22}{np.median(ts[3:]):8.1f}")

    def adam_tail():
        # the bench's default step: grad check -> unscale + Adam + EMA on the tail stream (bench.Variant, kind "replicated")
        from uda_poseestimation_b200.hotpath import ReplicatedTail
        stu = B.ParamBag(S.parameter_list(shapes, 1234 + 6, device=dev))
        tea = B.ParamBag([p.detach().clone() for p in stu.parameters()])
        tail = ReplicatedTail(stu, U.OldWeightEMA(tea, stu, alpha=0.999), algo="adam", loss_scale=B.LOSS_SCALE, lr=B.LR)
        for p, g in zip(stu.parameters(), B.synthetic_grads(shapes, 1234 + 9)):
            p.grad.copy_(g.to(dev))
        return dict(tail=tail, loss_scale=B.LOSS_SCALE), tea, stu

    if os.environ.get("PROBE_TIMELINE", "1") == "1":
        timeline("full step")
        timeline("EMA serial after the join", ema_parallel=False)
        kw, tea2, stu2 = adam_tail()
        timeline("Adam tail in the step (teacher chain as one launch)", models=(tea2, stu2), fuse_teacher_decode=True, **kw)

    def variant(name, ema=True, rewarp=True, skip=(), ema_parallel=True):
        inp = StepInputs(feat_src=d["feat_src"], feat_tgt_ori=d["feat_tgt_ori"], feat_tgt_tea=d["feat_tgt_tea"],
                         feat_src_ori=d["feat_src_ori"], y_s=d["y_s"], y_t_stu=d["y_t_stu"], y_t_tea=d["y_t_tea"],
                         label_s=label, weight_s=weight, alpha_s2t=alpha[0:1], alpha_t2s=alpha[1:2],
                         theta_tea=t_tea.to(dev) if rewarp else None, theta_stu=t_stu.to(dev) if rewarp else None)
        step = HotPathStep(teacher, student, sigma=cfg["sigma"], ema_parallel=ema_parallel)
        step.skip = frozenset(skip)
        step.capture(inp, include_ema=ema, warmup=2)
        for _ in range(5):
            step.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            step.replay()
        b.record()
        torch.cuda.synchronize()
        print(f"{name:<52}{a.elapsed_time(b) / args.iters * 1e3:9.1f} us", flush=True)

    from uda_poseestimation_b200.ema import MultiTensorPlan
    plan = MultiTensorPlan([p.data for p in teacher.parameters()], [p.data for p in student.parameters()])
    plan.run(0.999, 0.001, 0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        plan.run(0.999, 0.001, 0)
    b.record()
    torch.cuda.synchronize()
    print(f"EMA alone (eager launches)                          {a.elapsed_time(b) / 20 * 1e3:9.1f} us")
    variant("full step")
    variant("no EMA", ema=False)
    variant("no re-warp", rewarp=False)
    variant("no AdaIN", skip=("adain",))
    variant("no AdaIN, no EMA (heatmap chains alone)", ema=False, skip=("adain",))
    variant("no AdaIN, no EMA, no re-warp", ema=False, rewarp=False, skip=("adain",))
    variant("EMA serial after the join", ema_parallel=False)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-operator: the reference's eager PyTorch op sequence on CUDA tensors of THIS GPU (the oracle port functions,
which keep the reference's op order, Python loops, `.item()` syncs and D2H copies) against the package's operator,
same inputs, microbench sizes (C5: 256 x 21 x 64 x 64 heatmaps, 32 x 512 x 32 x 32 features, the PoseResNet-101
census) — SURVEY.md §2 "the bar for every hot-path row is the reference's eager op sequence on the same B200".

    python tools/eager_table.py [--out gpurun_out/eager_table] [--config C5]

Wall time between two device synchronisations (the reference's host work is part of what it costs), median of the
repetitions; writes <out>.json and <out>.md.  oracle/ is used here as the timed baseline, never by the product.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import uda_poseestimation_b200 as U  # noqa: E402
from oracle import reference_live, reference_port  # noqa: E402

# the reference's own functions where they can be loaded (tree, or oracle/_ref bytecode on the GPU box), else the port
R = reference_live if reference_live.available() else reference_port
EAGER = ("the reference's own functions (oracle/reference_live.py, " + str(reference_live.source()) + ")") if R is reference_live else "oracle/reference_port.py"
from uda_poseestimation_b200 import synthetic as S  # noqa: E402


def wall(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/eager_table")
    ap.add_argument("--config", default="C5")
    ap.add_argument("--recon-batch", type=int, default=32, help="batch of the re-warp rows (the reference loops per sample)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = S.CONFIGS[args.config]
    b, k, sigma = cfg["batch"], cfg["joints"], cfg["sigma"]
    rows = []

    def row(name, ref_site, shape, ref_fn, new_fn, ref_reps=3, new_reps=20, note=""):
        try:
            t_ref = wall(ref_fn, ref_reps)
        except Exception as exc:  # a reference path that cannot run on CUDA tensors is reported, not hidden
            t_ref, note = None, f"reference failed on CUDA: {type(exc).__name__}: {exc}"
        t_new = wall(new_fn, new_reps)
        rows.append(dict(operator=name, reference=ref_site, shape=shape, eager_us=t_ref, kernel_us=t_new,
                         speedup=(t_ref / t_new) if t_ref else None, note=note))
        print(f"{name:<34}{shape:<24} eager {t_ref if t_ref is None else round(t_ref, 1)!s:>12} us   kernel {t_new:9.1f} us   "
              f"x{(t_ref / t_new) if t_ref else float('nan'):8.1f}  {note}", flush=True)

    # features
    c, s_ = S.vgg_features(32, seed=1)
    c, s_ = c.to(dev), s_.to(dev)
    row("calc_mean_std", "adain/function.py:3-11", "32x512x32x32 f32", lambda: R.calc_mean_std(c), lambda: U.calc_mean_std(c))
    row("adain + alpha mix", "function.py:14-22, Style_net.py:167-168", "32x512x32x32 f32", lambda: R.adain_mix(c, s_, 0.37),
        lambda: U.adain_mix(c, s_, 0.37))
    c2, s2 = S.vgg_features(32, seed=2)
    c2, s2 = c2.to(dev), s2.to(dev)
    row("adain + mix, s2t and t2s", "train_human.py:348-356", "2 x 32x512x32x32 f32",
        lambda: (R.adain_mix(c, s_, 0.37), R.adain_mix(c2, s2, 0.8)), lambda: U.adain_mix_multi([(c, s_, 0.37), (c2, s2, 0.8)]))
    del c, s_, c2, s2
    # heatmaps
    shape = f"{b}x{k}x64x64"
    tea = S.heatmaps(b, k, seed=3, peak=(0.3, 1.2)).to(dev)
    stu = S.heatmaps(b, k, seed=4).to(dev).half()
    joints, vis = S.keypoints(b, k, seed=5)
    label, weight = U.generate_target_batched(joints, vis, (64, 64), sigma, (256, 256), device=dev)
    row("get_max_preds_torch", "utils.py:54-75", shape + " f32", lambda: R.get_max_preds_torch(tea), lambda: U.get_max_preds_torch(tea))
    row("rectify", "utils.py:77-109", shape + " f32", lambda: R.rectify(tea, sigma), lambda: U.rectify(tea, sigma), ref_reps=1,
        note="reference: B*K Python iterations with host syncs")
    row("conf / position / kth-mask", "train_human.py:376-383,427-430", shape + " f32",
        lambda: (R.confidence_mask(tea, 0.9), R.consistency_mask(tea, 0.5)),
        lambda: U.teacher_targets(tea, sigma, 0.5, occlude_thresh=0.9, materialise=False))
    row("accuracy (PCK)", "lib/keypoint_detection.py:65-94 via train_human.py:443", shape + " f16 / f32",
        lambda: R.accuracy(stu.detach().cpu().numpy(), label.detach().cpu().numpy()), lambda: U.accuracy(stu, label), ref_reps=2,
        note="reference: D2H of both tensors + numpy loops")

    def ref_mse():
        o = stu.detach().float().requires_grad_(True)      # autocast computes mse_loss in fp32
        (R.joints_mse_loss(o, label, weight) * 65536.0).backward()

    def new_mse():
        o = stu.detach().requires_grad_(True)
        (U.joints_mse_loss(o, label, weight) * 65536.0).backward()

    row("JointsMSELoss fwd+bwd", "lib/models/loss.py:39-49", shape + " f16 / f32", ref_mse, new_mse)
    rect = U.rectify(tea, sigma)
    mask = U.teacher_targets(tea, sigma, 0.5, materialise=False)["tea_mask"]

    def ref_cons():
        o = stu.detach().float().requires_grad_(True)
        (R.cons_loss(o, rect, tea_mask=mask) * 65536.0).backward()

    def new_cons():
        o = stu.detach().requires_grad_(True)
        (U.cons_loss(o, rect, tea_mask=mask) * 65536.0).backward()

    row("ConsLoss fwd+bwd", "lib/models/loss.py:124-132", shape + " f16 / f32", ref_cons, new_cons)
    row("generate_target (batch)", "lib/datasets/util.py:12-70 per sample", shape + " f32",
        lambda: [R.generate_target(joints[i], vis[i], (64, 64), sigma, (256, 256)) for i in range(b)],
        lambda: U.generate_target_batched(joints, vis, (64, 64), sigma, (256, 256), device=dev), ref_reps=2,
        note="reference: numpy on the host, per sample (loader workers)")
    sets = [S.keypoints(b, k, seed=60 + i)[0] for i in range(3)]
    row("generate_target x5 per sample", "rendered_hand_pose_mt.py:99-147", shape + " (3) + 8x8 (2)",
        lambda: R.loader_targets_hand(sets[0], sets[1], sets[2], vis, (64, 64), sigma, (256, 256)),
        lambda: U.generate_targets_multi([sets[0], sets[1], sets[0], sets[2], sets[2]], vis, [(64, 64), (64, 64), (8, 8), (64, 64), (8, 8)],
                                         sigma, (256, 256), device=dev), ref_reps=1, note="reference: numpy on the host")
    # re-warp loops (per-sample tF.affine): the trainers' batch
    rb = args.recon_batch
    tea_r = S.heatmaps(rb, k, seed=7, peak=(0.3, 1.2)).to(dev)
    stu_r = S.heatmaps(rb, k, seed=8).to(dev).half()
    aug_t, aug_s = S.aug_params(rb, seed=9), S.aug_params(rb, seed=10)
    rshape = f"{rb}x{k}x64x64"
    row("teacher re-warp (3 x tF.affine)", "train_human.py:359-372", rshape + " f32", lambda: R.teacher_recon([tea_r], [aug_t], 4.0),
        lambda: U.teacher_recon([tea_r], [aug_t], 4.0), ref_reps=2, note="reference: per-sample loop, CPU staging tensor")

    from uda_poseestimation_b200 import rewarp as RW
    theta_t = RW.stage_table(RW.recon_stages(aug_t, 4.0, rb), 64, 64, torch.float32, None)[0].to(dev)

    def ref_chain():     # :359-372 then :376-383 and :427-430 on the re-warped map
        rec = R.teacher_recon([tea_r], [aug_t], 4.0)
        R.confidence_mask(rec, 0.9)
        R.consistency_mask(rec, 0.5)

    row("teacher chain: re-warp -> conf / position / kth-mask", "train_human.py:359-383,427-430", rshape + " f32", ref_chain,
        lambda: U.teacher_targets_rewarped(tea_r, theta_t, sigma, 0.5, occlude_thresh=0.9), ref_reps=2,
        note="this package: ONE launch, the re-warped map is never written")

    def ref_stu():
        y = stu_r.detach().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.float16):
            out = R.student_recon(y, aug_s, 4.0, autocast=False)
        out.float().sum().backward()

    def new_stu():
        y = stu_r.detach().requires_grad_(True)
        U.student_recon(y, aug_s, 4.0, autocast=torch.float16).float().sum().backward()

    row("student re-warp fwd+bwd", "train_human.py:417-423", rshape + " f16", ref_stu, new_stu, ref_reps=2,
        note="reference: per-sample loop under autocast")
    # parameters
    shapes = S.pose_resnet_param_shapes(k)
    student = [torch.nn.Parameter(t) for t in S.parameter_list(shapes, 1, device=dev)]
    teacher = [torch.nn.Parameter(t.detach().clone()) for t in student]
    n_params = sum(p.numel() for p in student)
    pshape = f"PoseResNet-101 {n_params}"

    class Bag(torch.nn.Module):
        def __init__(self, ps):
            super().__init__()
            self.ps = torch.nn.ParameterList(ps)

    stu_m, tea_m = Bag(student), Bag(teacher)
    ema = U.OldWeightEMA(tea_m, stu_m, alpha=0.999)
    row("OldWeightEMA.step", "utils.py:21-25", pshape, lambda: R.ema_step([p.data for p in teacher], [p.data for p in student], 0.999),
        ema.step, note="reference: 969 launches")
    for p in student:
        p.grad = torch.randn_like(p) * 65.536
    ref_opt = torch.optim.Adam(student, lr=1e-4)
    ref_scaler = torch.amp.GradScaler("cuda", init_scale=65536.0, growth_interval=10 ** 9)
    ref_scaler.scale(torch.zeros((), device=dev))      # instantiates the scale tensor, as scaler.scale(loss) does

    def ref_tail():
        grads = [p.grad.clone() for p in student]      # backward rewrites the gradients every step; unscale_ is in place
        for p, g in zip(student, grads):
            p.grad = g
        ref_scaler.step(ref_opt)                        # unscale_ + found_inf .item() + foreach Adam
        R.ema_step([p.data for p in teacher], [p.data for p in student], 0.999)
        ref_scaler.update()

    new_opt = U.Adam(student, lr=1e-4)
    new_opt.attach_teacher(ema)
    scale = torch.full((), 65536.0, device=dev)

    def new_tail():
        new_opt.grad_scale, new_opt.found_inf = scale, new_opt.check_grads()
        new_opt.step()
        ema.step()

    row("scaler.step(Adam) + EMA", "train_human.py:436-440", pshape, ref_tail, new_tail, note="reference incl. a 212 MB gradient clone per step")
    out = Path(args.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    meta = dict(gpu=torch.cuda.get_device_name(0), torch=torch.__version__, config=args.config,
                how=f"wall time between device synchronisations, median; eager = {EAGER} on CUDA tensors")
    out.with_suffix(".json").write_text(json.dumps(dict(meta=meta, rows=rows), indent=1))
    lines = [f"# Reference eager op sequences vs the package's operators on one {meta['gpu']} ({args.config} sizes)", "",
             meta["how"] + ".", "", "| operator | reference site | shape | eager PyTorch (us) | this package (us) | speed-up | note |", "|---|---|---|---:|---:|---:|---|"]
    for r in rows:
        e = "failed" if r["eager_us"] is None else f"{r['eager_us']:.1f}"
        sp = "-" if r["speedup"] is None else f"{r['speedup']:.1f}x"
        lines.append(f"| {r['operator']} | `{r['reference']}` | {r['shape']} | {e} | {r['kernel_us']:.1f} | {sp} | {r['note']} |")
    out.with_suffix(".md").write_text("\n".join(lines) + "\n")
    print(f"wrote {out.with_suffix('.json')} and {out.with_suffix('.md')}")


if __name__ == "__main__":
    main()

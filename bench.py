#!/usr/bin/env python
"""Hot-path benchmark (driver contract: see README / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C2]

A *step* is one pass of the mean-teacher hot path (uda_poseestimation_b200.hotpath) over one
synthetic batch shard: AdaIN+mix s2t and t2s, re-warp of the teacher / student heatmaps to the
un-augmented frame (with the student's backward), teacher decode/masks/rectify, JointsMSELoss and
ConsLoss fwd+bwd, PCK counts, EMA over a PoseResNet-101-shaped parameter list.  At N=1 the
workload is BASELINE.json configs[1] (SURREAL->LSP, 16 keypoints, batch 32); at N>1 every
rank runs the same per-GPU batch (weak scaling) and the int32 PCK counts are all-reduced
over NCCL inside the timed region.

One JSON line is printed by rank 0:
  value      whole-job images/s with inputs resident in HBM (CUDA-graph replay + EMA launch)
  e2e        the same step through the public API with HOST (pinned) inputs: H2D of every
             input + D2H of losses / PCK counts / predictions inside the timed region
  roofline   the dominant kernel (EMA, ~56 % of the step's bytes) timed with CUDA events around
             each of its launches inside the timed region, vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference's CPU path timed on this box's host cores
`--impl reference` times that CPU path alone (the reference is pure Python on torch/numpy and
/root/reference does not exist on the GPU box, so the pinned oracle port is what runs).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from uda_poseestimation_b200 import synthetic as S  # noqa: E402

METRIC = "hot_path_images_per_sec"
UNIT = "images/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


# ------------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------------
def make_host_inputs(cfg: dict, seed: int, student_dtype=torch.float16, pin: bool = True) -> dict:
    """Seeded CPU tensors of one step (SURVEY.md §8d recipe)."""
    b, k, sigma = cfg["batch"], cfg["joints"], cfg["sigma"]
    src, tgt_style = S.vgg_features(b, seed)
    tgt, src_style = S.vgg_features(b, seed + 1)
    y_s = S.heatmaps(b, k, seed + 2, peak=(0.2, 1.1)).to(student_dtype)
    y_t_stu = S.heatmaps(b, k, seed + 3, peak=(0.2, 1.1)).to(student_dtype)
    y_t_tea = S.heatmaps(b, k, seed + 4, peak=(0.3, 1.2))
    joints, vis = S.keypoints(b, k, seed + 5)
    host = dict(feat_src=src, feat_tgt_ori=tgt_style, feat_tgt_tea=tgt, feat_src_ori=src_style, y_s=y_s,
                y_t_stu=y_t_stu, y_t_tea=y_t_tea, joints=torch.from_numpy(joints), vis=torch.from_numpy(vis))
    if pin and torch.cuda.is_available():
        host = {n: t.pin_memory() for n, t in host.items()}
    # meta['aug_param_tea'] / ['aug_param_stu'] of the batch, as the DataLoader collates them
    host["aug_tea"], host["aug_stu"] = S.aug_params(b, seed + 7), S.aug_params(b, seed + 8)
    return host


def stage_tables(host, student_dtype):
    """Host half of the re-warp: the per-sample tF.affine matrices -> two float32 [B,3,6] tables."""
    from uda_poseestimation_b200 import rewarp as RW

    b = host["y_t_tea"].shape[0]
    ac = student_dtype if student_dtype != torch.float32 else None
    t_tea = RW.stage_table(RW.recon_stages(host["aug_tea"], 4.0, b), 64, 64, torch.float32, None)[0]
    t_stu = RW.stage_table(RW.recon_stages(host["aug_stu"], 4.0, b), 64, 64, student_dtype, ac)[0]
    return t_tea, t_stu


class ParamBag(torch.nn.Module):
    def __init__(self, tensors):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(t) for t in tensors])


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md 'clocks' line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_ms: int = 100):
        self.index, self.period_ms, self.proc, self.lines = index, period_ms, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms),
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(self.period_ms / 1000.0 * 1.5)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference) — the reported baseline and the `--impl reference` arm
# ------------------------------------------------------------------------------------------------
def cpu_step_factory(cfg: dict, seed: int):
    """Returns (step_fn, batch): one reference-style CPU step on seeded inputs (fp32: the CPU has
    no fp16 autocast path; everything else as train_human.py:347-444)."""
    from oracle import reference_port as R  # test infrastructure, allowed here as the timed baseline

    host = make_host_inputs(cfg, seed, student_dtype=torch.float32, pin=False)
    b, k, sigma = cfg["batch"], cfg["joints"], cfg["sigma"]
    joints, vis = host["joints"].numpy(), host["vis"].numpy()
    lab = [R.generate_target(joints[i], vis[i], (64, 64), sigma, (256, 256)) for i in range(b)]
    label = torch.from_numpy(np.stack([x[0] for x in lab]))
    weight = torch.from_numpy(np.stack([x[1] for x in lab]))
    shapes = S.pose_resnet_param_shapes(k)
    student = S.parameter_list(shapes, seed + 6)
    teacher = [t.clone() for t in student]
    rng = np.random.RandomState(seed)

    def step():
        with torch.no_grad():
            t1 = R.adain_mix(host["feat_src"], host["feat_tgt_ori"], float(rng.uniform(0, 1)))
            t2 = R.adain_mix(host["feat_tgt_tea"], host["feat_src_ori"], float(rng.uniform(0, 1)))
            y_t_tea = R.teacher_recon([host["y_t_tea"]], [host["aug_tea"]], 4.0)   # train_human.py:359-372
            conf, pos, table = R.confidence_mask(y_t_tea, 0.9)
            mask, thresh, act = R.consistency_mask(y_t_tea, 0.5)
            rect = R.rectify(y_t_tea, sigma)
        y_s = host["y_s"].detach().requires_grad_(True)
        y_t = host["y_t_stu"].detach().requires_grad_(True)
        y_t_recon = R.student_recon(y_t, host["aug_stu"], 4.0)                     # :417-423
        loss = R.joints_mse_loss(y_s, label, weight) + 1.0 * R.cons_loss(y_t_recon, rect, tea_mask=mask)
        (loss * 65536.0).backward()
        R.ema_step(teacher, student, 0.999)
        acc, avg, cnt, pred = R.accuracy(y_s.detach().numpy(), label.numpy())
        return float(loss), avg, t1, t2, table

    return step, b


def time_cpu_path(cfg: dict, seed: int, steps: int, warmup: int, budget_s: float) -> dict:
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, b = cpu_step_factory(cfg, seed)
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    for _ in range(max(0, warmup - 1)):
        step()
    n = steps
    if budget_s is not None:
        n = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    mean = float(np.mean(ts))
    return dict(value=b / mean, unit=UNIT, cores=cores, kind="port", ms_per_step=mean * 1e3, steps=n,
                sample=f"{n} full steps of {cfg['name']} (batch {b}, {cfg['joints']} keypoints, fp32) after "
                       f"{warmup} warm-up; oracle/reference_port.py with torch.set_num_threads({cores})")


def run_reference_arm(args, cfg, rank):
    """The reference's own CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return
    r = time_cpu_path(cfg, 1234, args.steps, args.warmup, budget_s=None if args.steps <= 40 else 240.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args.gpus, graph=False),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json_line(line)


def workload_config(cfg, n_gpus, graph=True, fused=True, ema="graph"):
    return {"workload": f"{cfg['name']}: mean-teacher hot-path step (AdaIN s2t+t2s mix on 512x32x32 relu4_1 features, "
                        f"three-stage affine re-warp of teacher/student heatmaps (+ backward), "
                        f"teacher decode/conf/kth-mask/rectify, JointsMSE+Cons fwd+bwd, PCK, EMA over PoseResNet-101 "
                        f"params), batch {cfg['batch']}/GPU, {cfg['joints']} keypoints, 256x256 images, 64x64 heatmaps",
            "batch_per_gpu": cfg["batch"], "global_batch": cfg["batch"] * n_gpus, "keypoints": cfg["joints"],
            "sigma": cfg["sigma"], "parallelism": f"dp{n_gpus} (batch sharded, int32 PCK-count allreduce)",
            "l2": "inputs larger than L2: ~1.1 GB streamed per step (the 636 MB EMA pass evicts the 126 MB L2 "
                  "between steps)", "cuda_graph": graph,
            "losses": "fused loss step (one launch, teacher map evaluated from the arg-max)" if fused else
                      "operator by operator (fwd+bwd launches, materialised rectified map)",
            "ema_launch": ema}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200_arm(args, cfg, rank, world, local):
    import torch.distributed as dist

    import uda_poseestimation_b200 as U
    from uda_poseestimation_b200 import dist as D
    from uda_poseestimation_b200.hotpath import HotPathStep, StepInputs, step_algorithmic_bytes

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    U.load_library()
    numa_cpus = None
    if world > 1 and os.environ.get("UDAPE_NUMA_BIND", "1") == "1":
        # before any pinned allocation: this rank's host buffers belong in the memory next to its GPU
        numa_cpus = D.bind_to_gpu_numa(local)
    seed = 1234 + rank
    b, k, sigma = cfg["batch"], cfg["joints"], cfg["sigma"]
    host = make_host_inputs(cfg, seed)
    d = {n: t.to(dev, non_blocking=True) for n, t in host.items() if torch.is_tensor(t)}
    t_tea, t_stu = stage_tables(host, torch.float16)
    host["theta_tea"], host["theta_stu"] = t_tea.pin_memory(), t_stu.pin_memory()
    label, weight = U.generate_target_batched(d["joints"], d["vis"], (64, 64), sigma, (256, 256), device=dev)
    host["label_s"], host["weight_s"] = label.cpu().pin_memory(), weight.cpu().pin_memory()
    inp = StepInputs(feat_src=d["feat_src"], feat_tgt_ori=d["feat_tgt_ori"], feat_tgt_tea=d["feat_tgt_tea"],
                     feat_src_ori=d["feat_src_ori"], y_s=d["y_s"], y_t_stu=d["y_t_stu"], y_t_tea=d["y_t_tea"],
                     label_s=label, weight_s=weight, alpha_s2t=None, alpha_t2s=None,
                     theta_tea=host["theta_tea"].to(dev), theta_stu=host["theta_stu"].to(dev))
    alpha_pair = torch.zeros(2, device=dev)  # one 8-byte copy per step sets both directions
    inp.alpha_s2t, inp.alpha_t2s = alpha_pair[0:1], alpha_pair[1:2]
    shapes = S.pose_resnet_param_shapes(k)
    student = ParamBag(S.parameter_list(shapes, seed + 6, device=dev))
    teacher = ParamBag([torch.empty_like(p) for p in student.parameters()])
    ema_in_graph = args.ema in ("graph", "graph-serial") and not args.no_graph
    step = HotPathStep(teacher, student, sigma=sigma, fused=not args.unfused, ema_parallel=args.ema == "graph")
    n_steps_total = args.warmup + args.steps
    rng = np.random.RandomState(seed)  # alpha ~ U(0,1) per step (train_human.py:349,354)
    alphas = torch.from_numpy(rng.uniform(0, 1, size=(2 * n_steps_total + 64, 2)).astype(np.float32)).to(dev)

    use_graph = not args.no_graph
    # alpha ~ U(0,1) per step (train_human.py:349,354).  With the graph, the row of a device table is loaded for
    # the NEXT replay by three tiny torch kernels behind the AdaIN launches (a device step counter picks it), so a
    # replay needs no launch in front of it; eagerly, one 8-byte copy in front of the step
    alpha_row = torch.ones(1, dtype=torch.int64, device=dev)

    def alpha_feed():
        alpha_pair.copy_(alphas.index_select(0, alpha_row).view(2))
        alpha_row.add_(1).remainder_(alphas.shape[0])

    def set_alpha(i):
        if not use_graph:
            alpha_pair.copy_(alphas[i])

    if use_graph:
        step.alpha_feed = alpha_feed
    alpha_pair.copy_(alphas[0])
    # the path's only per-step exchange: int32 [2,K] PCK counts, summed over ranks.  It is issued on the
    # PCK chain (inside the graph when NCCL capture works) so that it overlaps the AdaIN / EMA chains.
    ar_in_step = False
    if world > 1:
        for _ in range(2):
            D.allreduce_counts(torch.zeros((2, k), dtype=torch.int32, device=dev))  # warm NCCL up before capture
        torch.cuda.synchronize()
        step.counts_hook = D.allreduce_counts
        ar_in_step = True
    if os.environ.get("UDAPE_BENCH_DEBUG") == "1":
        step.marks = []
    if use_graph:
        try:
            out = step.capture(inp, include_ema=ema_in_graph, warmup=2)
        except Exception as exc:  # NCCL not capturable on this stack: keep the all-reduce outside the graph
            if not ar_in_step:
                raise
            print(f"[bench] rank {rank}: graph capture with the NCCL all-reduce failed ({type(exc).__name__}: {exc}); "
                  "re-capturing without it", file=sys.stderr)
            torch.cuda.synchronize()
            step.counts_hook, ar_in_step = None, False
            out = step.capture(inp, include_ema=ema_in_graph, warmup=2)
        body = step.replay
    else:
        out = None
        body = lambda: step.run_no_ema(inp)  # noqa: E731

    def one_step(i, ev=None):
        set_alpha(i)
        o = body()
        if not ema_in_graph:
            if ev is not None:
                ev[0].record()
            step.ema.step()
            if ev is not None:
                ev[1].record()
        if world > 1 and not ar_in_step:
            D.allreduce_counts(o["pck_counts"])
        return o

    # ---- value: inputs resident in HBM -----------------------------------------------------------
    for i in range(args.warmup):
        one_step(i)
    ema_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start.record()
    for i in range(args.steps):
        out = one_step(args.warmup + i, ema_events[i])
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = start.elapsed_time(end)
    if os.environ.get("UDAPE_BENCH_DEBUG") == "1" and use_graph:
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a_.record()
        for _ in range(args.steps):
            step.replay()
        b_.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        acc = {}
        for _ in range(20):
            step.replay()
            torch.cuda.synchronize()
            for name, ev in step.marks[1:]:
                acc.setdefault(name, []).append(step.marks[0][1].elapsed_time(ev) * 1e3)
        for name, ts in sorted(acc.items(), key=lambda kv: np.median(kv[1])):
            print(f"[debug]   {name:<22}{np.median(ts):8.1f}", file=sys.stderr)
        print(f"[debug] tight replay loop: {a_.elapsed_time(b_) / args.steps * 1e3:.1f} us/step on the device, "
              f"{(t1 - t0) / args.steps * 1e6:.1f} us/step of host launch time; timed loop {ms_total / args.steps * 1e3:.1f} us/step",
              file=sys.stderr)
    if ema_in_graph:
        # the EMA kernel runs inside the graph (concurrently with the other chains when --ema graph), where
        # it cannot be bracketed by events: time the same launch on its own right after the step loop
        # (parameters >> L2, so every launch streams from HBM)
        for a, b_ in ema_events:
            a.record()
            step.ema.step()
            b_.record()
        torch.cuda.synchronize()
    ema_ms = float(np.mean([a.elapsed_time(b_) for a, b_ in ema_events]))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * b / (ms_per_step / 1e3)

    # ---- e2e: host (pinned) inputs -> device -> step -> results back on the host -----------------
    # the source labels travel as keypoints (12 KB) and become heatmaps on the device (udape_gauss_target, row
    # a4): the reference builds them in its loader workers and ships 8.4 MB of float32 heatmaps per step
    h2d_names = ["feat_src", "feat_tgt_ori", "feat_tgt_tea", "feat_src_ori", "y_s", "y_t_stu", "y_t_tea",
                 "theta_tea", "theta_stu"]
    h2d_bytes = sum(host[n].numel() * host[n].element_size() for n in h2d_names + ["joints", "vis"]) + 8
    res_host = dict(losses=torch.empty(3, dtype=torch.float32).pin_memory(),
                    counts=torch.empty((2, k), dtype=torch.int32).pin_memory(),
                    pred=torch.empty((b, k, 2), dtype=torch.float32).pin_memory())
    d2h_bytes = sum(t_.numel() * t_.element_size() for t_ in res_host.values())
    alpha_host = torch.from_numpy(rng.uniform(0, 1, size=(n_steps_total, 2)).astype(np.float32)).pin_memory()
    losses_dev = torch.empty(3, dtype=torch.float32, device=dev)

    def e2e_step(i):
        for n in h2d_names[:-2]:
            getattr(inp, n).copy_(host[n], non_blocking=True)
        d["joints"].copy_(host["joints"], non_blocking=True)
        d["vis"].copy_(host["vis"], non_blocking=True)
        U.generate_target_batched(d["joints"], d["vis"], (64, 64), sigma, (256, 256), out=(inp.label_s, inp.weight_s))
        # host half of the re-warp (the reference computes the same matrices inside tF.affine, per sample);
        # it runs while the asynchronous copies above are on the wire
        t_tea, t_stu = stage_tables(host, torch.float16)
        host["theta_tea"].copy_(t_tea)
        host["theta_stu"].copy_(t_stu)
        for n in h2d_names[-2:]:
            getattr(inp, n).copy_(host[n], non_blocking=True)
        alpha_pair.copy_(alpha_host[i], non_blocking=True)   # this step's alpha comes from the host
        o = body()
        if not ema_in_graph:
            step.ema.step()
        if world > 1 and not ar_in_step:
            D.allreduce_counts(o["pck_counts"])
        torch.stack((o["loss_all"], o["loss_s"], o["loss_c"]), out=losses_dev)
        res_host["losses"].copy_(losses_dev, non_blocking=True)
        res_host["counts"].copy_(o["pck_counts"], non_blocking=True)
        res_host["pred"].copy_(o["pred"], non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads loss / accuracy every step
        acc, avg, cnt = U.accuracy_from_counts(res_host["counts"][0], res_host["counts"][1])
        return float(res_host["losses"][0]), avg

    e2e_warm = max(1, min(args.warmup, 3))
    for i in range(e2e_warm):
        e2e_step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    start.record()
    for i in range(args.steps):
        last = e2e_step(i)
    end.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(start.elapsed_time(end), wall_ms)  # host work is part of end-to-end
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t.item()) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- optional: the gradient all-reduce measured on its own (N>1) ------------------------------
    grad_ar = None
    if world > 1:
        flat = torch.zeros(step.n_params, dtype=torch.float32, device=dev)
        for _ in range(3):
            dist.all_reduce(flat)
        torch.cuda.synchronize()
        start.record()
        reps = 10
        for _ in range(reps):
            dist.all_reduce(flat)
        end.record()
        torch.cuda.synchronize()
        t = torch.tensor([start.elapsed_time(end) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ar_ms = float(t.item())
        nbytes = flat.numel() * 4
        grad_ar = {"ms": ar_ms, "bytes": nbytes, "busbw_GBps": 2 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9,
                   "note": "student-gradient allreduce (fp32 flat bucket), timed on its own; in training it overlaps "
                           "the cuDNN backward and is therefore not part of the hot-path step"}

    if rank != 0:
        return
    # ---- roofline + baselines ------------------------------------------------------------------
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "B200_PROFILING.md fallback (of fallback)"
    abytes = step_algorithmic_bytes(inp, step.n_params, fused=step.fused)
    achieved = abytes["ema"] / (ema_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "ema_traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "ema_multi_kernel<float> (udape_ema_multi)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": abytes["ema"], "avg_launch_ms": ema_ms,
                "timed": ("CUDA events around each of the K EMA launches inside the timed step loop" if not ema_in_graph else
                          "CUDA events around K EMA launches issued right after the timed step loop (inside it the "
                          "kernel is a CUDA-graph node overlapping the other chains and cannot be bracketed)"),
                "step_algorithmic_bytes": abytes["total"],
                "step_frac_of_peak": abytes["total"] / (ms_per_step * 1e-3) / 1e9 / peak}
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        r = time_cpu_path(cfg, 1234, steps=args.cpu_steps, warmup=1, budget_s=25.0)
        cpu = {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (fp16 student heatmaps, fp32 accumulation)", "data": "synthetic",
        "config": workload_config(cfg, world, graph=use_graph, fused=step.fused,
                                  ema=args.ema if use_graph else "after"),
        "clocks": clocks,
        "e2e": {"value": world * b / (e2e_ms_per_step / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms_per_step,
                "last_loss": last[0], "last_avg_pck": last[1],
                "labels": "keypoints copied per step, heatmaps generated on the device (udape_gauss_target)",
                "host_numa_binding": (f"rank 0 pinned to {len(numa_cpus)} CPUs local to its GPU (NVML affinity)"
                                      if numa_cpus else "none")},
        "gpu_launches": args.steps * step.kernels_per_step,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if grad_ar is not None:
        line["grad_allreduce"] = grad_ar
        line["config"]["pck_allreduce"] = "inside the step graph (PCK chain)" if (ar_in_step and use_graph) else "after the step"
    emit_json_line(line)


_JSON_FD = None


def claim_stdout():
    """The driver parses stdout as ONE JSON line.  Native libraries write there too (NCCL prints its
    version banner on fd 1 at NCCL_DEBUG=VERSION and above), so fd 1 is pointed at stderr for the whole
    run and the JSON line is written to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json_line(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", choices=sorted(S.CONFIGS), default="C2")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--unfused", action="store_true", help="operator-by-operator losses (separate fwd/bwd launches, "
                    "materialised rectified teacher map) instead of the fused loss step")
    ap.add_argument("--ema", choices=["graph", "graph-serial", "after"], default="graph",
                    help="where the EMA launch sits: a parallel branch of the step graph (default), the last node of "
                         "the graph, or a separate launch after the graph replay")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=8)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: >= 3 warm-up steps
    cfg = S.CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    relaunch = args.impl == "b200" and args.gpus > 1 and "WORLD_SIZE" not in os.environ
    if not relaunch:
        claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args, cfg, rank)  # rank 0 alone works; the others exit 0
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference for the CPU path)")
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself as one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), str(Path(__file__).resolve()),
               *sys.argv[1:]]
        raise SystemExit(subprocess.call(cmd))
    if world > 1:
        # keep stdout to the single JSON line: NCCL's version banner (NCCL_DEBUG=VERSION) goes to stdout
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        from uda_poseestimation_b200 import dist as D
        D.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    try:
        run_b200_arm(args, cfg, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()

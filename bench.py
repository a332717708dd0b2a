#!/usr/bin/env python
"""Hot-path benchmark (driver contract: see README / DESIGN.md §5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C2] [--tail auto|peer|nccl|ema]

A *step* is one pass of the mean-teacher hot path (uda_poseestimation_b200.hotpath) over one synthetic batch
shard — everything train_human.py:347-444 does between the cuDNN passes: AdaIN+mix s2t and t2s, re-warp of the
teacher / student heatmaps (with the student's backward), teacher decode / masks / rectify, JointsMSELoss and
ConsLoss fwd+bwd, PCK counts, and the tail of the step  scaler.step(stu_optimizer); tea_optimizer.step()
(:436-438): the data-parallel gradient exchange, GradScaler's non-finite check + unscale, Adam, and the teacher
EMA over a PoseResNet-101-shaped parameter list.  At N=1 the workload is BASELINE.json configs[1] (SURREAL->LSP,
16 keypoints, batch 32); at N>1 every rank runs the same per-GPU batch (weak scaling), the student gradients and
the int32 PCK counts are exchanged INSIDE the timed step.

One JSON line is printed by rank 0:
  value       whole-job images/s with inputs resident in HBM (CUDA-graph replays over rotating input sets)
  variants    the same step with the other tails: "nccl" (NCCL all-reduce -> replicated update), "ema" (round-1
              definition: bare EMA, no gradient exchange / optimizer), so the cost of the exchange is visible
  e2e         the same step through the public API with HOST (pinned) inputs: H2D of every input (double-
              buffered against the previous step) + D2H of losses / PCK counts / predictions, in the timed region
  roofline    the dominant kernel (student_step at N=1; the peer-memory gather+EMA at N>1) timed with CUDA
              events on its own launches, vs MEASURED_PEAKS.json; `nvlink` at N>1: bytes pulled over NVLink / time
  cpu_baseline / eager_cuda_baseline   the reference's own functions (oracle/_ref bytecode; the restated port when absent) on this box's host cores / the
              same functions as eager PyTorch on CUDA tensors of the same GPU (incl. its host syncs and copies)
  multi_gpu_parity (N>1)  N-rank results == single-process results on the rank-ordered sum (tools/dp_parity.py)
`--impl reference` times the CPU path alone: the reference is pure Python on torch/numpy; /root/reference does not
exist on the GPU box, but the bytecode oracle/build_ref.py compiled from its hot-path files (oracle/_ref/, outputs
only) does, so what runs is the reference's own functions and objects with the trainer's inline loops restated
around them (oracle/reference_live.py; `cpu_baseline.kind` says which: "reference", or "port" without the bytecode).
"""
from __future__ import annotations

import argparse
import contextlib
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
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback
NVLINK_PEER_GBS = 770.0        # B200_PROFILING.md: measured peer copy, per direction per GPU (900 nominal)
LOSS_SCALE = 65536.0           # GradScaler's initial scale (train_human.py:324)
LR = 1e-4                      # train_human.py --lr default for Adam
GRAD_STD = 0.01                # unscaled gradient magnitude of the synthetic bucket


# ------------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------------
def make_host_inputs(cfg: dict, seed: int, student_dtype=torch.float16, pin: bool = True) -> dict:
    """Seeded CPU tensors of one step (SURVEY.md §8d recipe)."""
    b, k = cfg["batch"], cfg["joints"]
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


def synthetic_grads(shapes, seed: int, device="cpu"):
    """Loss-scaled gradient of every parameter (what `scaler.scale(loss_all).backward()` leaves in p.grad)."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    return [(torch.randn(*s, generator=g) * (GRAD_STD * LOSS_SCALE)).to(device) for s in shapes]


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
# The reference's path as the reference runs it: oracle port on CPU tensors (the reported baseline and the
# `--impl reference` arm) or on CUDA tensors (eager PyTorch on the same GPU: what each kernel replaces)
# ------------------------------------------------------------------------------------------------
def reference_step_factory(cfg: dict, seed: int, device="cpu", tail: bool = True):
    """Returns (step_fn, batch, kind): one reference-style step on seeded inputs, everything as train_human.py:347-444
    (fp32 on the CPU, which has no fp16 autocast path; fp16 student maps on CUDA), `tail`: incl. GradScaler
    unscale + Adam (:436-437) in front of the EMA (:438).

    kind "reference": the reference's OWN functions and objects (oracle/reference_live.py: imported from the
    reference tree, or from the bytecode oracle/build_ref.py compiled from it into oracle/_ref/, which travels to
    the GPU box) with the trainer's inline loops restated around them; kind "port": oracle/reference_port.py
    alone, when neither is present."""
    from oracle import reference_live as RL  # test infrastructure, allowed here as the timed baseline
    from oracle import reference_port as RP

    live = RL.available()
    R = RL if live else RP
    dev = torch.device(device)
    on_gpu = dev.type == "cuda"
    host = make_host_inputs(cfg, seed, student_dtype=torch.float16 if on_gpu else torch.float32, pin=False)
    b, k, sigma = cfg["batch"], cfg["joints"], cfg["sigma"]
    joints, vis = host["joints"].numpy(), host["vis"].numpy()
    lab = [R.generate_target(joints[i], vis[i], (64, 64), sigma, (256, 256)) for i in range(b)]
    label = torch.from_numpy(np.stack([x[0] for x in lab])).to(dev)
    weight = torch.from_numpy(np.stack([x[1] for x in lab])).to(dev)
    t = {n: (v.to(dev) if torch.is_tensor(v) else v) for n, v in host.items()}
    shapes = S.pose_resnet_param_shapes(k)
    student = S.parameter_list(shapes, seed + 6, device=dev)
    teacher = [x.clone() for x in student]
    grads = synthetic_grads(shapes, seed + 9, device=dev) if tail else None
    trainer_tail = ema = None
    if live:
        # the reference's own objects: torch.optim.Adam + GradScaler + OldWeightEMA (train_human.py:139-141, :324)
        trainer_tail = RL.TrainerTail(student, teacher, LR, 0.999, LOSS_SCALE, dev)
        scaler, ema = trainer_tail.scaler, trainer_tail.ema
    else:
        m = [torch.zeros_like(x) for x in student] if tail else None
        v = [torch.zeros_like(x) for x in student] if tail else None
    state = {"step": 0}
    rng = np.random.RandomState(seed)

    def step():
        with torch.no_grad():
            t1 = R.adain_mix(t["feat_src"], t["feat_tgt_ori"], float(rng.uniform(0, 1)))
            t2 = R.adain_mix(t["feat_tgt_tea"], t["feat_src_ori"], float(rng.uniform(0, 1)))
            y_t_tea = R.teacher_recon([t["y_t_tea"]], [host["aug_tea"]], 4.0)          # train_human.py:359-372
            conf, pos, table = R.confidence_mask(y_t_tea, 0.9)
            mask, thresh, act = R.consistency_mask(y_t_tea, 0.5)
            rect = R.rectify(y_t_tea, sigma)
        y_s = t["y_s"].detach().requires_grad_(True)
        y_t = t["y_t_stu"].detach().requires_grad_(True)
        with (torch.autocast("cuda", dtype=torch.float16) if on_gpu else contextlib.nullcontext()):     # :414
            y_t_recon = R.student_recon(y_t, host["aug_stu"], 4.0, autocast=not on_gpu)  # :417-423
            loss = R.joints_mse_loss(y_s, label, weight) + 1.0 * R.cons_loss(y_t_recon, rect, tea_mask=mask)
        if live:
            scaler.scale(loss).backward()                                               # :435
            if tail:
                trainer_tail.step(grads)                                                # :436-441
            else:
                ema.step()                                                              # :438
                scaler.update()
        else:
            (loss * LOSS_SCALE).backward()
            if tail:
                _, state["step"] = R.student_teacher_step("adam", student, [g.clone() for g in grads], m, v, teacher,
                                                          state["step"], LOSS_SCALE, 0.999, lr=LR)   # :436-438
            else:
                R.ema_step(teacher, student, 0.999)
        acc, avg, cnt, pred = R.accuracy(y_s.detach().cpu().numpy(), label.cpu().numpy())   # :443-444
        return float(loss), avg, t1, t2, table

    return step, b, ("reference" if live else "port")


def time_reference_path(cfg: dict, seed: int, steps: int, warmup: int, budget_s: float | None, device="cpu", tail: bool = True) -> dict:
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, b, kind = reference_step_factory(cfg, seed, device=device, tail=tail)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    t0 = time.perf_counter()
    step()
    sync()
    first = time.perf_counter() - t0
    for _ in range(max(0, warmup - 1)):
        step()
    sync()
    n = steps
    if budget_s is not None:
        n = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        step()
        sync()
        ts.append(time.perf_counter() - t0)
    mean, med = float(np.mean(ts)), float(np.median(ts))
    from oracle import ref_loader
    impl = ("the reference's own functions and objects (adain, rectify, JointsMSELoss, ConsLoss, accuracy, generate_target, "
            f"OldWeightEMA, torch Adam + GradScaler; loaded from {'the reference tree' if ref_loader.kind() == 'source' else 'oracle/_ref bytecode compiled from the reference tree'}) "
            "with the trainer's inline re-warp loops / masks restated around them (oracle/reference_live.py)"
            if kind == "reference" else "oracle/reference_port.py (restated; reference files not present on this box)")
    what = (f"{n} full steps of {cfg['name']} (batch {b}, {cfg['joints']} keypoints) after {warmup} warm-up, incl. "
            f"{'GradScaler unscale + Adam + ' if tail else ''}EMA over the PoseResNet-101 census; {impl}")
    if device == "cpu":
        return dict(value=b / mean, unit=UNIT, cores=cores, kind=kind, ms_per_step=mean * 1e3, steps=n,
                    sample=what + f", fp32, torch.set_num_threads({cores})")
    return dict(value=b / med, unit=UNIT, ms_per_step=med * 1e3, steps=n, kind=kind,
                sample=what + " as eager PyTorch on CUDA tensors of this GPU (fp16 student maps; its per-sample "
                              "tF.affine loops, CPU staging tensors, .item() syncs and the D2H for accuracy() included); median")


def run_reference_arm(args, cfg, rank):
    """The reference's own CPU implementation of the path on this box's host cores (rank 0 alone: the reference
    is one process; at N>1 the line is labelled with what actually ran — one process, one batch per step)."""
    if rank != 0:
        return
    r = time_reference_path(cfg, 1234, args.steps, args.warmup, budget_s=None if args.steps <= 40 else 240.0,
                            tail=args.tail != "ema")
    conf = workload_config(cfg, 1, graph=False, tail="cpu")
    conf["parallelism"] = "one CPU process (the reference's nn.DataParallel is single-process); one batch per step"
    conf["requested_gpus"] = args.gpus
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": conf,
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json_line(line)


def workload_config(cfg, n_gpus, graph=True, fused=True, tail="peer", sets=2):
    tails = {"peer": "gradient reduce-scatter + non-finite check + Adam on the rank's 1/N slice (one kernel) -> parameter "
                     "all-gather + EMA (one kernel), peer loads over NVLink, no NCCL in the step",
             "nccl": "NCCL all-reduce of the flat fp32 gradient bucket -> grad check -> unscale + Adam + EMA (one launch)",
             "replicated": "grad check -> unscale + Adam + EMA in one multi-tensor launch",
             "ema": "bare EMA (round-1 step definition: no gradient exchange, no optimizer)",
             "cpu": "GradScaler unscale + torch-style Adam + EMA on the host"}
    return {"workload": f"{cfg['name']}: mean-teacher hot-path step (AdaIN s2t+t2s mix on 512x32x32 relu4_1 features, "
                        f"three-stage affine re-warp of teacher/student heatmaps (+ backward), "
                        f"teacher decode/conf/kth-mask/rectify, JointsMSE+Cons fwd+bwd, PCK, and the step's tail over the "
                        f"PoseResNet-101 parameter census), batch {cfg['batch']}/GPU, {cfg['joints']} keypoints, 256x256 "
                        f"images, 64x64 heatmaps",
            "batch_per_gpu": cfg["batch"], "global_batch": cfg["batch"] * n_gpus, "keypoints": cfg["joints"],
            "sigma": cfg["sigma"], "parallelism": f"dp{n_gpus} (batch sharded; student-gradient and int32 PCK-count "
                                                  f"exchange inside the step)",
            "tail": tails.get(tail, tail),
            "l2": f"inputs larger than L2 and rotated: {sets} input sets alternate between replays and every step streams "
                  ">= 1.1 GB (126 MB L2)", "cuda_graph": graph,
            "losses": "fused loss step (one launch, teacher map evaluated from the arg-max)" if fused else
                      "operator by operator (fwd+bwd launches, materialised rectified map)"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Variant:
    """One assembled step (HotPathStep + tail) with its own parameters / optimizer state, captured once per
    rotating input set."""

    def __init__(self, kind, args, cfg, rank, world, dev, inputs, feed):
        import uda_poseestimation_b200 as U
        from uda_poseestimation_b200 import dist as D
        from uda_poseestimation_b200 import dp as DP
        from uda_poseestimation_b200.hotpath import HotPathStep, PeerTail, ReplicatedTail

        self.kind, self.world, self.dev = kind, world, dev
        k = cfg["joints"]
        shapes = S.pose_resnet_param_shapes(k)
        seed = 1234 + rank
        # the student starts identical on every rank (one model, replicated); gradients differ per rank
        student = ParamBag(S.parameter_list(shapes, 1234 + 6, device=dev))
        teacher = ParamBag([p.detach().clone() for p in student.parameters()])
        self.student, self.teacher = student, teacher
        self.peers = None
        tail = None
        if kind == "peer":
            _, n_total = DP.flat_layout(list(student.parameters()))
            self.peers = U.PeerGroup.create(DP.arena_bytes(n_total), dev)
            tail = PeerTail(student, teacher, self.peers, algo="adam", alpha=0.999, loss_scale=LOSS_SCALE, lr=LR,
                            timeout_s=60.0, capturable=False)
        elif kind in ("nccl", "replicated"):
            ema = U.OldWeightEMA(teacher, student, alpha=0.999)
            tail = ReplicatedTail(student, ema, algo="adam", loss_scale=LOSS_SCALE, lr=LR)
        if tail is not None:
            for p, g in zip(student.parameters(), synthetic_grads(shapes, seed + 9)):
                p.grad.copy_(g.to(dev))
        self.tail = tail
        hook = None
        if world > 1:
            hook = None if kind == "peer" else D.allreduce_counts   # the peer tail exchanges the counts itself
        # (one teacher view + the fused loss step: re-warp, decode, conf_table and the k-th value mask of the teacher chain
        # are ONE launch and the re-warped map is never written)
        self.step = HotPathStep(teacher, student, sigma=cfg["sigma"], fused=not args.unfused, tail=tail, counts_hook=hook,
                                loss_scale=LOSS_SCALE, fuse_teacher_decode=not args.unfused)
        self.step.alpha_feed = feed
        self.graphs = []
        self.inputs = inputs
        self.use_graph = not args.no_graph
        if self.use_graph:
            for inp in inputs:
                out = self.step.capture(inp, include_ema=True, warmup=2)
                self.graphs.append((self.step.graph, out))
        self.n = 0

    def run(self, i=None):
        i = self.n if i is None else i
        self.n += 1
        s = i % len(self.inputs)
        if self.use_graph:
            g, out = self.graphs[s]
            g.replay()
            return out
        return self.step.run(self.inputs[s])

    def close(self):
        self.graphs.clear()
        self.step = None
        if self.peers is not None:
            self.tail = None
            self.peers.close()


def timed_loop(variant, steps, warmup, world, dev):
    import torch.distributed as dist
    for _ in range(warmup):
        variant.run()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start.record()
    for _ in range(steps):
        out = variant.run()
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps, out


def run_b200_arm(args, cfg, rank, world, local):
    import torch.distributed as dist

    import uda_poseestimation_b200 as U
    from uda_poseestimation_b200 import dist as D
    from uda_poseestimation_b200.hotpath import ScalarFeed, StepInputs, step_algorithmic_bytes

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    U.load_library()
    numa_cpus = None
    if world > 1 and os.environ.get("UDAPE_NUMA_BIND", "1") == "1":
        numa_cpus = D.bind_to_gpu_numa(local)   # before any pinned allocation
    seed = 1234 + rank
    b, k, sigma = cfg["batch"], cfg["joints"], cfg["sigma"]
    n_sets = max(1, args.sets)
    rng = np.random.RandomState(seed)  # alpha ~ U(0,1) per step (train_human.py:349,354)
    n_steps_total = args.warmup + args.steps
    # alpha of every step: a device table; ONE single-warp launch of this library behind the AdaIN launches loads
    # the next step's pair (no host copy, no framework kernels inside the step graph)
    feed = ScalarFeed(torch.from_numpy(rng.uniform(0, 1, size=(2 * n_steps_total + 64, 2)).astype(np.float32)).to(dev))
    hosts, inputs, dev_kv = [], [], []
    for s_ in range(n_sets):
        host = make_host_inputs(cfg, seed + 100 * s_)
        d = {n: t.to(dev, non_blocking=True) for n, t in host.items() if torch.is_tensor(t)}
        t_tea, t_stu = stage_tables(host, torch.float16)
        host["theta_tea"], host["theta_stu"] = t_tea.pin_memory(), t_stu.pin_memory()
        label, weight = U.generate_target_batched(d["joints"], d["vis"], (64, 64), sigma, (256, 256), device=dev)
        inputs.append(StepInputs(feat_src=d["feat_src"], feat_tgt_ori=d["feat_tgt_ori"], feat_tgt_tea=d["feat_tgt_tea"],
                                 feat_src_ori=d["feat_src_ori"], y_s=d["y_s"], y_t_stu=d["y_t_stu"], y_t_tea=d["y_t_tea"],
                                 label_s=label, weight_s=weight, alpha_s2t=feed.out[0:1], alpha_t2s=feed.out[1:2],
                                 theta_tea=host["theta_tea"].to(dev), theta_stu=host["theta_stu"].to(dev)))
        hosts.append(host)
        dev_kv.append(d)
    torch.cuda.synchronize()

    # ---- which tail ------------------------------------------------------------------------------
    kind = args.tail
    notes = []
    if kind == "auto":
        kind = "peer" if world > 1 else "replicated"
    if kind == "nccl" and world == 1:
        kind = "replicated"
    if world > 1:
        for _ in range(2):   # warm NCCL up before any capture (PCK exchange of the nccl variant, parity, timing reduce)
            D.allreduce_counts(torch.zeros((2, k), dtype=torch.int32, device=dev))
        torch.cuda.synchronize()

    def build(kind_):
        if world > 1:
            dist.barrier()
        v = Variant(kind_, args, cfg, rank, world, dev, inputs, feed)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return v

    try:
        main = build(kind)
    except Exception as exc:
        if kind != "peer":
            raise
        # peer memory could not be set up on this box (no IPC / no P2P): the NCCL tail is the same arithmetic
        notes.append(f"peer tail unavailable ({type(exc).__name__}: {exc}); NCCL tail used")
        print(f"[bench] rank {rank}: {notes[-1]}", file=sys.stderr)
        kind = "nccl"
        main = build(kind)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_per_step, out = timed_loop(main, args.steps, args.warmup, world, dev)
    value = world * b / (ms_per_step / 1e3)
    if main.peers is not None:
        main.tail.opt.check()

    # ---- e2e: host (pinned) inputs -> device -> step -> results back on the host -----------------
    e2e = run_e2e(args, cfg, main, hosts, inputs, dev_kv, feed, rank, world, dev, rng)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the dominant kernel on its own launches + NVLink figures --------------------------------
    abytes = step_algorithmic_bytes(inputs[0], main.step.n_params, fused=main.step.fused, tail=main.tail,
                                    fuse_teacher_decode=main.step.fuse_teacher_decode)
    probe = probe_tail(main, world, dev, reps=max(5, min(args.steps, 20)))

    # ---- the other tails, same inputs ------------------------------------------------------------
    variants = {kind: {"ms_per_step": ms_per_step, "value": value, "step_algorithmic_bytes": abytes["total"]}}
    alts = [] if args.no_variants else (["nccl", "ema"] if world > 1 else ["ema"])
    for alt in alts:
        if alt == kind:
            continue
        try:
            v = build(alt)
            ms, _ = timed_loop(v, min(args.steps, 50), max(3, min(args.warmup, 5)), world, dev)
            ab = step_algorithmic_bytes(inputs[0], v.step.n_params, fused=v.step.fused, tail=v.tail,
                                            fuse_teacher_decode=v.step.fuse_teacher_decode)
            variants[alt] = {"ms_per_step": ms, "value": world * b / (ms / 1e3), "step_algorithmic_bytes": ab["total"]}
            if alt == "nccl":
                variants[alt]["grad_allreduce"] = time_nccl_allreduce(v, world, dev)
            v.close()
            del v
        except Exception as exc:  # a variant is context, never the headline
            variants[alt] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    parity = parity_detail = None
    if world > 1 and not args.no_parity:
        sys.path.insert(0, str(ROOT / "tools"))
        import dp_parity
        try:
            res = dp_parity.check(dev, keypoints=k, batch=b)
            parity = "ok" if all(res.values()) else "FAILED"
            parity_detail = res
        except Exception as exc:
            parity, parity_detail = "error", f"{type(exc).__name__}: {exc}"

    if rank != 0:
        main.close()
        return
    # ---- roofline + baselines (rank 0) -----------------------------------------------------------
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "B200_PROFILING.md fallback (of fallback)"
    dom = probe["dominant"]
    achieved = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "kernel_traffic.json"
    if tf.exists():
        try:
            rec = json.loads(tf.read_text()).get(dom["traffic_key"])
            if rec:
                traffic, traffic_src = rec["dram_bytes_per_launch"], rec["source"] + " (ncu capture, not this run)"
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom["bytes"], "avg_launch_ms": dom["ms"],
                "timed": dom["timed"], "tail_kernels_ms": probe["kernels_ms"],
                "step_algorithmic_bytes": abytes["total"], "step_bytes_by_kernel": {n: v for n, v in abytes.items() if n != "total"},
                "step_frac_of_peak": abytes["total"] / (ms_per_step * 1e-3) / 1e9 / peak,
                "step_frac_note": "algorithmic HBM bytes of the step / step time / peak; the few-MB tensors the heatmap chains "
                                  "hand from kernel to kernel are meant to stay in L2, so this counts L2-served bytes too"}
    if probe.get("nvlink"):
        roofline["nvlink"] = probe["nvlink"]
    cpu = eager = None
    if world == 1 and not args.skip_cpu_baseline:
        r = time_reference_path(cfg, 1234, steps=args.cpu_steps, warmup=1, budget_s=25.0, tail=kind != "ema")
        cpu = {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
        try:
            main.close()
            torch.cuda.empty_cache()
            r = time_reference_path(cfg, 1234, steps=5, warmup=2, budget_s=60.0, device=str(dev), tail=kind != "ema")
            eager = {k_: r[k_] for k_ in ("value", "unit", "ms_per_step", "steps", "kind", "sample")}
        except Exception as exc:
            eager = {"error": f"{type(exc).__name__}: {exc}"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (fp16 student heatmaps, fp32 accumulation)", "data": "synthetic",
        "config": workload_config(cfg, world, graph=main.use_graph, fused=not args.unfused, tail=kind, sets=n_sets),
        "clocks": clocks, "e2e": e2e,
        "gpu_launches": args.steps * probe["kernels_per_step"],
        "roofline": roofline, "cpu_baseline": cpu, "eager_cuda_baseline": eager,
        "variants": variants,
        "grad_allreduce": "in-step" if kind in ("peer", "nccl") else ("none (single GPU)" if kind == "replicated" else "separate / absent"),
    }
    if world > 1:
        line["multi_gpu_parity"] = parity
        line["multi_gpu_parity_detail"] = parity_detail
        line["config"]["pck_allreduce"] = "inside the step graph (PCK chain), " + ("peer-memory kernel" if kind == "peer" else "NCCL")
    if notes:
        line["notes"] = notes
    emit_json_line(line)


def probe_tail(main, world, dev, reps):
    """CUDA events around the tail's own launches, issued eagerly right after the timed loop (inside the step the
    kernels are graph nodes overlapping the other chains and cannot be bracketed); every rank issues the same
    sequence, so the cross-rank protocol of the peer tail runs as in the step.  Mean over `reps`."""
    import torch.distributed as dist
    step, tail = main.step, main.tail
    p4 = 4 * step.n_params
    # + the alpha table feed, + the PCK-count exchange kernel of the peer tail
    kernels_per_step = step.kernels_per_step + 1 + (1 if (world > 1 and tail is not None and tail.name == "peer") else 0)
    mk = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    if tail is None:
        evs = [(mk(), mk()) for _ in range(reps)]
        for a, b_ in evs:
            a.record()
            step.ema.step()
            b_.record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b_) for a, b_ in evs]))
        return {"dominant": {"kernel": "ema_multi_kernel<float> (udape_ema_multi)", "bytes": 3 * p4, "ms": ms,
                             "traffic_key": "ema_multi", "timed": f"CUDA events around {reps} launches right after the timed loop"},
                "kernels_ms": {"ema_multi": ms}, "kernels_per_step": kernels_per_step}
    if tail.name == "replicated":
        # each kernel as a one-node CUDA graph replayed back to back (no host gaps between the launches; the
        # 2.1 GB the update touches per launch is >> L2, so every replay streams from HBM)
        opt = tail.opt
        opt.grad_scale = tail.scale

        def graph_ms(fn):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            for _ in range(3):
                g.replay()
            a, b_ = mk(), mk()
            a.record()
            for _ in range(reps):
                g.replay()
            b_.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b_) / reps

        def check():
            opt.found_inf = opt.check_grads()
        check()
        kms = {"grad_check": graph_ms(check), "student_step": graph_ms(opt.step)}
        return {"dominant": {"kernel": "student_step_kernel<ADAM> (udape_student_step: unscale + Adam + teacher EMA)",
                             "bytes": 9 * p4, "ms": kms["student_step"], "traffic_key": "student_step_adam",
                             "timed": f"CUDA events around {reps} back-to-back replays of a one-kernel CUDA graph, right after the "
                                      "timed loop (parameters >> L2)"},
                "kernels_ms": kms, "kernels_per_step": kernels_per_step}
    # peer tail: events between its four launches
    opt = tail.opt
    names = ["barrier", "reduce_step", "wait_reduced", "gather_ema"]
    acc = {n: [] for n in names}
    if world > 1:
        dist.barrier()
    for _ in range(reps):
        evs = [mk() for _ in range(5)]
        opt.grad_scale = tail.scale
        opt.step(_events=evs)
        torch.cuda.synchronize()
        for i, n in enumerate(names):
            acc[n].append(evs[i].elapsed_time(evs[i + 1]))
    opt.check()
    kms = {n: float(np.mean(v)) for n, v in acc.items()}
    # max over ranks (a phase ends when the slowest rank's kernel ends)
    t = torch.tensor([kms[n] for n in names], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kms = {n: float(x) for n, x in zip(names, t.tolist())}
    bts = tail.bytes()
    w, s4 = world, 4 * opt.shard_elems
    pulled = (w - 1) * s4
    nv = {"bytes_in_per_kernel": pulled, "peak_GBps_per_direction": NVLINK_PEER_GBS,
          "peak_source": "B200_PROFILING.md measured peer copy, one direction (900 nominal); here every GPU pulls and "
                         "serves at once, so read requests share each link with the data flowing the other way",
          "reduce_step_GBps": pulled / (kms["reduce_step"] * 1e-3) / 1e9 if pulled else None,
          "gather_ema_GBps": pulled / (kms["gather_ema"] * 1e-3) / 1e9 if pulled else None}
    if pulled:
        nv["reduce_step_frac"] = nv["reduce_step_GBps"] / NVLINK_PEER_GBS
        nv["gather_ema_frac"] = nv["gather_ema_GBps"] / NVLINK_PEER_GBS
    dom = "gather_ema" if kms["gather_ema"] >= kms["reduce_step"] else "reduce_step"
    kernel = {"gather_ema": "dp_gather_ema_kernel (udape_dp_gather_ema: parameter all-gather by peer loads + teacher EMA)",
              "reduce_step": "dp_reduce_step_kernel (udape_dp_reduce_step: gradient reduce-scatter by peer loads + non-finite "
                             "check + unscale + Adam on the rank's slice)"}[dom]
    return {"dominant": {"kernel": kernel, "bytes": bts[dom], "ms": kms[dom], "traffic_key": "dp_" + dom,
                         "timed": f"CUDA events between the tail's launches, {reps} eager tails right after the timed loop, max over ranks"},
            "kernels_ms": kms, "nvlink": nv if world > 1 else None, "kernels_per_step": kernels_per_step}


def time_nccl_allreduce(v, world, dev, reps=10):
    import torch.distributed as dist
    flat = v.tail.bucket.flat
    keep = flat.clone()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        dist.all_reduce(flat)
    torch.cuda.synchronize()
    start.record()
    for _ in range(reps):
        dist.all_reduce(flat)
    end.record()
    torch.cuda.synchronize()
    flat.copy_(keep)
    t = torch.tensor([start.elapsed_time(end) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, nbytes = float(t.item()), flat.numel() * 4
    return {"ms": ms, "bytes": nbytes, "busbw_GBps": 2 * (world - 1) / world * nbytes / (ms * 1e-3) / 1e9,
            "note": "dist.all_reduce of the flat fp32 bucket on its own (what the 'nccl' variant has inside its step)"}


def run_e2e(args, cfg, main, hosts, inputs, dev_kv, feed, rank, world, dev, rng):
    """The step through the public API with HOST inputs.  Two input sets double-buffer the copies: while the
    graph of step i runs on set A, a copy stream moves step i+1's host inputs into set B (features, heatmaps,
    keypoints -> label heatmaps built on the device, re-warp tables, the gradient bucket); results of step i come
    back through pinned buffers and are read by the host every step (one step behind the launch)."""
    import torch.distributed as dist

    import uda_poseestimation_b200 as U
    b, k, sigma = cfg["batch"], cfg["joints"], cfg["sigma"]
    n_sets = len(inputs)
    names = ["feat_src", "feat_tgt_ori", "feat_tgt_tea", "feat_src_ori", "y_s", "y_t_stu", "y_t_tea"]
    grad_dev = grad_host = None
    if main.tail is not None:
        grad_dev = main.tail.opt.flat_grads if main.tail.name == "peer" else main.tail.bucket.flat
        grad_host = grad_dev.cpu().pin_memory()
    h2d_bytes = sum(hosts[0][n].numel() * hosts[0][n].element_size() for n in names + ["joints", "vis", "theta_tea", "theta_stu"]) + 8
    if grad_host is not None:
        h2d_bytes += grad_host.numel() * 4
    res_host = [dict(losses=torch.empty(3, dtype=torch.float32).pin_memory(),
                     counts=torch.empty((2, k), dtype=torch.int32).pin_memory(),
                     pred=torch.empty((b, k, 2), dtype=torch.float32).pin_memory()) for _ in range(n_sets)]
    d2h_bytes = sum(t_.numel() * t_.element_size() for t_ in res_host[0].values())
    losses_dev = [torch.empty(3, dtype=torch.float32, device=dev) for _ in range(n_sets)]
    alpha_host = torch.from_numpy(rng.uniform(0, 1, size=(args.steps + 8, 2)).astype(np.float32)).pin_memory()
    copy_stream = torch.cuda.Stream(dev)
    cur = torch.cuda.current_stream()
    copied = [torch.cuda.Event() for _ in range(n_sets)]      # set s holds the inputs of its next step
    consumed = [torch.cuda.Event() for _ in range(n_sets)]    # the step that read set s has finished
    done = [torch.cuda.Event() for _ in range(n_sets)]        # results of the step on set s are on the host
    for ev in consumed:
        ev.record(cur)

    def stage(i):
        """enqueue the H2D copies of step i's inputs into set i % n_sets (copy stream)"""
        s_ = i % n_sets
        host, inp, d = hosts[s_], inputs[s_], dev_kv[s_]
        # host half of the re-warp (the reference computes the same matrices inside tF.affine, per sample)
        t_tea, t_stu = stage_tables(host, torch.float16)
        host["theta_tea"].copy_(t_tea)
        host["theta_stu"].copy_(t_stu)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s_])
            for n in names:
                getattr(inp, n).copy_(host[n], non_blocking=True)
            d["joints"].copy_(host["joints"], non_blocking=True)
            d["vis"].copy_(host["vis"], non_blocking=True)
            inp.theta_tea.copy_(host["theta_tea"], non_blocking=True)
            inp.theta_stu.copy_(host["theta_stu"], non_blocking=True)
            # the source labels travel as keypoints (12 KB) and become heatmaps on the device (udape_gauss_target,
            # row a4): the reference builds them in its loader workers and ships 8.4 MB of float32 heatmaps per step
            U.generate_target_batched(d["joints"], d["vis"], (64, 64), sigma, (256, 256), out=(inp.label_s, inp.weight_s))
            copied[s_].record(copy_stream)

    def stage_grads():
        # the gradient bucket is one buffer (backward writes it in place every step): copied on the copy stream
        # behind the previous step's tail, which is the last reader
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(tail_done)
            grad_dev.copy_(grad_host, non_blocking=True)
            grads_ready.record(copy_stream)

    tail_done, grads_ready = torch.cuda.Event(), torch.cuda.Event()
    tail_done.record(cur)

    def launch(i):
        s_ = i % n_sets
        cur.wait_event(copied[s_])
        if grad_host is not None:
            cur.wait_event(grads_ready)
        feed.out.copy_(alpha_host[i % alpha_host.shape[0]], non_blocking=True)   # this step's alpha comes from the host
        o = main.run(i)
        consumed[s_].record(cur)
        tail_done.record(cur)
        torch.stack((o["loss_all"], o["loss_s"], o["loss_c"]), out=losses_dev[s_])
        res_host[s_]["losses"].copy_(losses_dev[s_], non_blocking=True)
        res_host[s_]["counts"].copy_(o["pck_counts"], non_blocking=True)
        res_host[s_]["pred"].copy_(o["pred"], non_blocking=True)
        done[s_].record(cur)

    def read(i):
        s_ = i % n_sets
        done[s_].synchronize()        # the caller reads loss / accuracy of every step
        acc, avg, cnt = U.accuracy_from_counts(res_host[s_]["counts"][0], res_host[s_]["counts"][1])
        return float(res_host[s_]["losses"][0]), avg

    def loop(n):
        last = None
        stage(0)
        if grad_host is not None:
            stage_grads()
        for i in range(n):
            launch(i)
            if i + 1 < n:
                stage(i + 1)
                if grad_host is not None:
                    stage_grads()
            if i > 0:
                last = read(i - 1)
        return read(n - 1)

    # the box's H2D ceiling for these very buffers (copies alone, all ranks at once)
    def h2d_only(n):
        for i in range(n):
            s_ = i % n_sets
            for nm in names:
                getattr(inputs[s_], nm).copy_(hosts[s_][nm], non_blocking=True)
            if grad_host is not None:
                grad_dev.copy_(grad_host, non_blocking=True)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d_only(2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start.record()
    h2d_only(6)
    end.record()
    torch.cuda.synchronize()
    big = sum(hosts[0][n].numel() * hosts[0][n].element_size() for n in names) + (grad_host.numel() * 4 if grad_host is not None else 0)
    t = torch.tensor([start.elapsed_time(end) / 6], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    copy_ms = float(t.item())
    ceiling = big / (copy_ms * 1e-3) / 1e9

    loop(max(2, min(args.warmup, 3)))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = args.steps
    t0 = time.perf_counter()
    start.record()
    last = loop(steps)
    end.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(start.elapsed_time(end), wall_ms)  # host work is part of end-to-end
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"value": world * b / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
            "ms_per_step": ms, "last_loss": last[0], "last_avg_pck": last[1],
            "pipeline": f"{n_sets} input sets: the copies of step i+1 overlap the graph of step i (copy stream + events); "
                        "results are read by the host every step",
            "h2d_copy_ms_per_step": copy_ms, "h2d_ceiling_GBps_per_gpu": ceiling,
            "h2d_ceiling_GBps_aggregate": ceiling * world,
            "h2d_ceiling_note": "the step's own pinned buffers copied back to back with no compute, all ranks at once (max over ranks)",
            "frac_of_copy_bound": copy_ms / ms,
            "labels": "keypoints copied per step, heatmaps generated on the device (udape_gauss_target)",
            "gradients": "the flat gradient bucket (what backward leaves) is a host input of the step too" if grad_host is not None else "none"}


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
    ap.add_argument("--tail", choices=["auto", "peer", "nccl", "replicated", "ema"], default="auto",
                    help="what follows backward inside the step: auto = peer-memory exchange + sharded Adam + EMA at N>1, "
                         "grad check + fused Adam + EMA at N=1; nccl = NCCL all-reduce + replicated update; ema = bare EMA "
                         "(round-1 step definition)")
    ap.add_argument("--sets", type=int, default=2, help="rotating input sets (>= 2: L2 cannot serve re-reads across replays)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--unfused", action="store_true", help="operator-by-operator losses (separate fwd/bwd launches, "
                    "materialised rectified teacher map) instead of the fused loss step")
    ap.add_argument("--no-variants", action="store_true", help="skip the other tails' timing")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank parity check (N>1)")
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

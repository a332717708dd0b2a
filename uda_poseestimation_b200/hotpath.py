"""The per-step hot-path sequence of the reference's ``train()`` (``train_human.py:347-444``)
assembled from the drop-in operators — everything the trainer does between the cuDNN
forward/backward passes, on tensors the convolutions would have produced:

    s2t / t2s   AdaIN + alpha mix of relu4_1 features        (:348-356 via Style_net.py:167-168)
    teacher     re-warp of the teacher heatmaps to the        (:359-372)
                un-augmented frame (three tF.affine stages)
    teacher     conf / position / conf_table, activates,      (:376-383, :427-430)
                rectify, k-th-value consistency mask
    student     re-warp of the student target heatmaps        (:417-423) + its backward
    student     JointsMSELoss fwd+bwd, ConsLoss fwd+bwd        (:425, :432, :434-436)
    teacher     EMA update over the PoseResNet parameter list  (:438)
    metric      PCK hit/valid counts on (y_s, label_s)         (:443-444)

The ResNet/VGG/decoder convolutions and the Adam step are out of scope (cuDNN / torch); their
outputs are the inputs of this step.  ``HotPathStep.run()`` makes only public-API calls, is
free of host synchronisation, and can therefore be captured into a CUDA graph
(``capture()``); alpha scalars live in device memory so each replay can use fresh values.
"""
from __future__ import annotations

import ctypes
import dataclasses
import os

import torch

from . import _lib
from . import rewarp as _rewarp
from .adain import adain_mix, adain_mix_multi
from .ema import OldWeightEMA
from .keypoint_detection import _pck
from .loss import cons_loss, fused_losses, joints_mse_loss
from .mask import teacher_targets, teacher_targets_rewarped

__all__ = ["StepInputs", "HotPathStep", "step_algorithmic_bytes", "ScalarFeed", "ReplicatedTail", "PeerTail"]


class ScalarFeed:
    """Per-step scalars of a captured step: row ``counter % rows`` of a device table is copied to ``out`` and
    the counter advances, in ONE single-warp launch (``udape_table_feed``) — alpha ~ U(0,1) per step
    (train_human.py:349,354) without a host copy or framework kernels in front of a graph replay."""

    def __init__(self, table: torch.Tensor):
        if table.dim() != 2 or table.dtype != torch.float32 or not table.is_contiguous():
            raise TypeError("ScalarFeed: contiguous float32 [rows, cols] table")
        self.dev = _lib.require_cuda(table)
        self.table = table
        self.out = table[0].clone()
        self.counter = torch.ones(1, dtype=torch.int32, device=self.dev)   # row 0 is loaded; the next load is row 1

    def __call__(self):
        t = self.table
        with _lib.on_device(self.dev):
            st = _lib.load().udape_table_feed(t.data_ptr(), t.shape[0], t.shape[1], self.counter.data_ptr(),
                                              self.out.data_ptr(), _lib.stream_ptr(self.dev))
        _lib.check(st, "udape_table_feed")


class ReplicatedTail:
    """``scaler.step(stu_optimizer); tea_optimizer.step()`` (train_human.py:436-438) with every rank holding the
    whole optimizer state: [NCCL all-reduce of the flat gradient bucket (world > 1)] -> ``udape_grad_check`` ->
    ``udape_student_step`` (unscale + Adam | SGD + teacher EMA in one multi-tensor launch)."""

    name = "replicated"

    def __init__(self, student, ema: OldWeightEMA, algo: str = "adam", loss_scale: float = 65536.0, group=None, **hyper):
        from . import dist as D
        from .optim import SGD, Adam

        self.world = D.dist.get_world_size(group) if (D.dist.is_available() and D.dist.is_initialized()) else 1
        self.group = group
        self.bucket = D.FlatGradBucket(list(student.parameters()))
        self.opt = (Adam if algo == "adam" else SGD)(student.parameters(), **hyper)
        self.opt.attach_teacher(ema)
        self.ema, self.algo = ema, algo
        self.n_params = self.bucket.flat.numel()
        self.scale = torch.full((), float(loss_scale), dtype=torch.float32, device=self.bucket.flat.device)
        self.kernels = 2   # grad_check + student_step (the NCCL kernel is the library's, not counted)
        self.mark = None   # profiling only: HotPathStep._mark of the step that records a timeline

    @property
    def found_inf(self):
        return self.opt._found_inf

    def begin(self):
        if self.world > 1:
            self.bucket.allreduce_(average=True, group=self.group)      # DataParallel's reduce-add, as a mean

    def finish(self, counts=None):
        # (the PCK counts of this tail travel by ncclAllReduce from the PCK chain: HotPathStep.counts_hook)
        self.opt.grad_scale, self.opt.found_inf = self.scale, self.opt.check_grads()
        if self.mark is not None:
            self.mark("grad check done")      # profiling only (tools/step_probe.py)
        self.opt.step()
        self.ema._fused_pending = False   # tea_optimizer.step() is not called separately by the assembled step

    def run(self):
        self.begin()
        self.finish()

    def bytes(self) -> dict:
        p4 = 4 * self.n_params
        passes = 9 if self.algo == "adam" else 7
        out = {"grad_check": p4, "student_step": passes * p4}
        if self.world > 1:
            out["allreduce_hbm (NCCL ring: ~2 reads + 2 writes of the bucket)"] = 4 * p4
        return out


class PeerTail:
    """The same tail as two peer-memory kernels over NVLink (``dp.ShardedStudentStep``): gradient reduce-scatter +
    non-finite check + the update of this rank's 1/W slice in one kernel, parameter all-gather + EMA in the other."""

    name = "peer"

    def __init__(self, student, teacher, peers, algo: str = "adam", alpha: float = 0.999, loss_scale: float = 65536.0, **hyper):
        from .dp import ShardedStudentStep

        self.opt = ShardedStudentStep(student.parameters(), peers, algo=algo, teacher_params=list(teacher.parameters()),
                                      alpha=alpha, **hyper)
        self.algo, self.world = algo, peers.world
        self.n_params = self.opt.n_total
        self.scale = torch.full((), float(loss_scale), dtype=torch.float32, device=peers.dev)
        self.kernels = self.opt.kernels_per_step

    @property
    def found_inf(self):
        return self.opt.found_inf

    exchanges_counts = True   # finish(counts) sums the PCK counts over the ranks on the tail's own stream

    def begin(self):
        self.opt.grad_scale = self.scale
        self.opt.step_begin()

    def finish(self, counts=None):
        self.opt.step_finish(counts if self.world > 1 else None)

    def run(self):
        self.begin()
        self.finish()

    def bytes(self) -> dict:
        w, p4, s4 = self.world, 4 * self.n_params, 4 * self.opt.shard_elems
        state = 6 if self.algo == "adam" else 4      # p, m(, v) read; p', m'(, v') written  (+ w gradient slices read)
        return {"reduce_step": (w + state) * s4,     # of which (w - 1) * s4 arrive over NVLink
                "gather_ema": p4 + p4 + 2 * p4,      # shadow read (NVLink but for the own slice), params written, teacher r+w
                "nvlink_in": 2 * (w - 1) * s4}


@dataclasses.dataclass
class StepInputs:
    """Device tensors one step consumes (what the convs / loader hand to the hot path)."""
    feat_src: torch.Tensor      # [N,512,32,32] relu4_1(x_s)            content of s2t
    feat_tgt_ori: torch.Tensor  # [N,512,32,32] relu4_1(x_t_teas_ori)   style of s2t
    feat_tgt_tea: torch.Tensor  # [N,512,32,32] relu4_1(x_t_tea)        content of t2s
    feat_src_ori: torch.Tensor  # [N,512,32,32] relu4_1(x_s_ori)        style of t2s
    y_s: torch.Tensor           # [B,K,64,64] student(x_s)      fp16 under autocast
    y_t_stu: torch.Tensor       # [B,K,64,64] student(x_t_stu)  fp16 under autocast (re-warped)
    y_t_tea: torch.Tensor       # [B,K,64,64] teacher(x_t_tea)  fp32 — or a list of k views (train_human.py:358,361:
                                # every view is re-warped with its own table and the views are averaged)
    label_s: torch.Tensor       # [B,K,64,64] fp32 Gaussian target
    weight_s: torch.Tensor      # [B,K,1]     fp32 visibility weight
    alpha_s2t: torch.Tensor     # [1] fp32 device scalar
    alpha_t2s: torch.Tensor     # [1] fp32 device scalar
    # re-warp stage tables (rewarp.stage_table of the batch's aug_param_tea / aug_param_stu, float32
    # [B,3,6] on the device, refreshed per step).  None: y_t_tea / y_t_stu are taken as already re-warped.
    theta_tea: torch.Tensor | None = None      # one table, or a list of k tables when y_t_tea is a list of views
    theta_stu: torch.Tensor | None = None
    # occlusion (train_human.py:385-412; eager runs only — it needs the host's np.random stream, like the reference):
    # the student's target images and the collated meta['aug_param_stu'] of the batch
    x_t_stu: torch.Tensor | None = None        # [B,3,256,256]
    aug_param_stu: list | None = None

    def tensors(self):
        out = []
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            if torch.is_tensor(v):
                out.append(v)
            elif isinstance(v, (list, tuple)) and v and all(torch.is_tensor(t) for t in v):
                out.extend(v)
        return out

    @property
    def teacher_views(self):
        return list(self.y_t_tea) if isinstance(self.y_t_tea, (list, tuple)) else [self.y_t_tea]


def step_algorithmic_bytes(inp: StepInputs, n_params: int, param_bytes: int = 4, fused: bool = True, tail=None,
                           fuse_teacher_decode: bool = False) -> dict:
    """Algorithmic HBM bytes of one step, per kernel family.

    ``fused=True`` is the default step (one loss launch, rectified teacher map evaluated on the fly);
    ``fused=False`` the operator-by-operator sequence (separate fwd/bwd launches, materialised map).
    ``tail``: the ReplicatedTail / PeerTail of the step (its kernels replace the bare EMA pass).
    ``fuse_teacher_decode``: the teacher re-warp is arg-maxed where it is gathered (one read, no map written or re-read)."""
    ef = inp.feat_src.element_size()
    feat = inp.feat_src.numel() * ef
    tea0 = inp.teacher_views[0]
    hm = tea0.numel()
    e_s, e_t, e_l = inp.y_s.element_size(), tea0.element_size(), inp.label_s.element_size()
    planes = inp.y_s.shape[0] * inp.y_s.shape[1]
    out = {
        "adain_mix": 2 * 3 * feat,                        # two directions x (2 reads + 1 write)
        "mask_select": 9 * planes,
        "pck": hm * (e_s + e_l) + 8 * planes,
    }
    if tail is None:
        out["ema"] = 3 * n_params * param_bytes
    else:
        for name, nbytes in tail.bytes().items():
            if name != "nvlink_in":      # bytes over NVLink are not HBM bytes of this rank (reported separately)
                out[name] = nbytes
    if fused:
        out["decode"] = hm * e_t + 40 * planes            # 1 read (+ per-plane outputs)
        # y_s + label read, grad_y_s written; y_t_stu read, grad_y_t_stu written (teacher map analytic)
        out["loss_step"] = hm * (e_s + e_l + e_s) + hm * (e_s + e_s) + 8 * planes
    else:
        out["decode_rectify"] = hm * e_t + hm * e_t + 40 * planes  # 1 read + 1 write (+ per-plane outputs)
        out["joints_mse_fwd"] = hm * (e_s + e_l) + 4 * planes
        out["joints_mse_bwd"] = hm * (e_s + e_l) + hm * e_s
        out["cons_fwd"] = hm * (e_s + e_t) + 4 * planes
        out["cons_bwd"] = hm * (e_s + e_t) + hm * e_s
    if inp.theta_tea is not None:
        k_views = len(inp.teacher_views)
        if fuse_teacher_decode and fused and k_views == 1:
            out["rewarp_teacher+decode"] = hm * e_t + 72 * tea0.shape[0] + 40 * planes    # one read; the map is never written
            del out["decode"]
        else:
            out["rewarp_teacher"] = (k_views + 1) * hm * e_t + 72 * k_views * tea0.shape[0]   # k reads + 1 write (+ the stage tables)
    if inp.theta_stu is not None:
        out["rewarp_student_fwd"] = 2 * hm * e_s + 72 * inp.y_t_stu.shape[0]
        out["rewarp_student_bwd"] = 2 * hm * e_s + 72 * inp.y_t_stu.shape[0]   # grad read + grad write
    out["total"] = sum(v for k, v in out.items())
    return out


class HotPathStep:
    """One mean-teacher hot-path step on a batch shard, through the public operators."""

    def __init__(self, teacher: torch.nn.Module, student: torch.nn.Module, sigma=2, mask_ratio: float = 0.5,
                 occlude_thresh: float = 0.9, teacher_alpha: float = 0.999, lambda_c: float = 1.0,
                 loss_scale: float = 65536.0, parallel: bool = True, fused: bool = True,
                 ema_parallel: bool = True, counts_hook=None, tail=None, ema: OldWeightEMA | None = None,
                 occlude_rate: float = 0.5, occlude_size: int = 10, image_size: int = 256, rng=None,
                 fuse_teacher_decode: bool = False):
        self.sigma, self.mask_ratio, self.occlude_thresh = sigma, mask_ratio, occlude_thresh
        self.lambda_c, self.loss_scale = lambda_c, loss_scale
        self.parallel, self.fused, self.ema_parallel = parallel, fused, ema_parallel
        # called with the int32 [2,K] hits||valid tensor right after the PCK launch, on the PCK chain's
        # stream: the data-parallel integer all-reduce (dist.allreduce_counts) goes here so that it
        # overlaps the AdaIN / EMA chains instead of trailing the step
        self.counts_hook = counts_hook
        self.occlude_rate, self.occlude_size, self.image_size = occlude_rate, occlude_size, image_size
        self.rng = rng          # np.random-like stream of the occlusion draws (None: numpy's global one, as the reference)
        # one teacher view + the fused loss step: the re-warped teacher map is only ever decoded, so the teacher chain is ONE
        # launch that arg-maxes the planes where it gathers them (rewarp.gather_decode) and out["y_t_tea_recon"] is None
        self.fuse_teacher_decode = fuse_teacher_decode
        self._side = None
        # tail: what follows backward — None: the bare EMA (:438, the student update left to torch);
        # ReplicatedTail / PeerTail: gradient exchange + unscale + Adam | SGD + EMA (:436-438)
        self.tail = tail
        self.ema = ema if ema is not None else (getattr(tail, "ema", None) or
                                                (OldWeightEMA(teacher, student, alpha=teacher_alpha) if tail is None else None))  # :141
        self.n_params = sum(p.numel() for p in teacher.parameters())
        self.graph = None
        self.out = None
        self._graph_inputs = None
        self.rewarp_kernels = 0
        # optional callable run on the AdaIN stream right AFTER the two AdaIN launches (inside the captured
        # graph): loads the device-resident alpha scalars of the NEXT replay, e.g. from a device table indexed
        # by a device step counter, so that a replay needs no separate launch in front of it.  (In front of the
        # AdaIN launches it would delay them behind the EMA's 13k CTAs: a lower-priority grid that has started
        # is not displaced, measured 188 -> 199 us.)
        self.alpha_feed = None
        self.adain_one_launch = os.environ.get("UDAPE_ADAIN_ONE_LAUNCH", "1") == "1"
        self.skip = frozenset()   # profiling only (tools/step_probe.py): chains left out of the step
        self.marks = None         # profiling only: list of (name, external timing event) filled while capturing

    def _mark(self, name: str):
        if self.marks is not None:
            ev = torch.cuda.Event(enable_timing=True, external=True)
            ev.record(torch.cuda.current_stream())
            self.marks.append((name, ev))

    @property
    def kernels_per_step(self) -> int:
        """Kernels of libudape_b200.so launched by run() (memset nodes for tickets/counters not counted):
        fused: 2 adain, decode (+ k-th select in its last CTA), loss_step, pck, ema;  unfused: 2 adain,
        decode+rectify(+select), mse fwd/bwd, cons fwd/bwd, pck, ema;  + 4 with the re-warp tables (teacher forward,
        student forward, its inverse plan, student backward)."""
        return ((6 if self.fused else 9) + self.rewarp_kernels + (self.tail.kernels - 1 if self.tail is not None else 0)
                - (1 if self.adain_one_launch else 0))

    def _streams(self, dev):
        if self._side is None or self._side[0].device != dev:
            # the heatmap chains are short, latency-bound kernels on the step's critical path: give them
            # priority over the bandwidth-bound AdaIN / EMA launches when CTAs compete for SM slots
            # ... and the two AdaIN passes priority over the EMA, which then fills whatever the others leave
            # (measured on B200, tools/step_probe.py: 214 -> 199 us per step; equal priorities make the
            # block scheduler drain the EMA's 13k CTAs before the first AdaIN CTA is dispatched)
            on = os.environ.get("UDAPE_STEP_PRIORITY", "1") == "1"
            hi = int(os.environ.get("UDAPE_CHAIN_PRIORITY", "-2")) if on else 0
            ad = int(os.environ.get("UDAPE_ADAIN_PRIORITY", "-1")) if on else 0
            tl = int(os.environ.get("UDAPE_TAIL_PRIORITY", "0")) if on else 0
            self._side = (torch.cuda.Stream(dev, priority=hi), torch.cuda.Stream(dev, priority=hi), torch.cuda.Stream(dev, priority=tl),
                          torch.cuda.Stream(dev, priority=hi), torch.cuda.Stream(dev, priority=ad))
        return self._side

    # -- the step -----------------------------------------------------------------------------------
    def _run(self, inp: StepInputs, with_ema: bool) -> dict:
        """Independent chains, forked onto side streams so that the small heatmap kernels
        (launch/latency-bound at batch 32) and the EMA stream overlap the two large AdaIN passes
        (stream priority in brackets):

            adain  [-1]: AdaIN+mix s2t, AdaIN+mix t2s [, alpha_feed for the next replay]
            teacher[-2]: [re-warp ->] decode + k-th select (one launch) -> fused loss step (both criteria +
                         both gradients) [-> re-warp backward of the consistency gradient]
            student[-2]: [re-warp of y_t_stu ->] PCK counts   [unfused: JointsMSELoss fwd -> bwd -> PCK]
            plan   [-2]: [inverse plan of the student re-warp, consumed by its backward]
            ema    [ 0]: multi-tensor EMA over the parameter list (after the join if ema_parallel=False)

        The fork/join is plain stream-event ordering, so it behaves the same eagerly and under
        CUDA-graph capture (where it becomes parallel graph branches)."""
        cur = torch.cuda.current_stream()
        s_tea, s_stu, s_ema, s_plan, s_adain = self._streams(inp.y_s.device) if self.parallel else (cur, cur, cur, cur, cur)

        ema_side = with_ema and self.parallel and self.ema_parallel
        self._mark("start")
        if self.parallel:
            s_tea.wait_stream(cur)
            s_stu.wait_stream(cur)
            if ema_side:
                s_ema.wait_stream(cur)
        if ema_side:
            with torch.cuda.stream(s_ema):
                # :438 — independent of every other chain of the hot path (in training it follows
                # scaler.step(stu_optimizer); the student parameters are an input of this step)
                if self.tail is not None:
                    self.tail.begin()      # :436 — the gradient exchange starts with the step
                else:
                    self.ema.step()
                    self._mark("ema done")
        # teacher forward; student forward + inverse plan + backward
        tea_one_launch = (self.fuse_teacher_decode and self.fused and inp.theta_tea is not None and len(inp.teacher_views) == 1
                          and _rewarp.gather_decode_supported(inp.teacher_views[0]))
        self.rewarp_kernels = ((1 if inp.theta_tea is not None else 0)
                               + ((3 if _rewarp.USE_INVERSE_PLAN else 2) if inp.theta_stu is not None else 0)
                               + (1 if inp.x_t_stu is not None else 0)
                               - (1 if tea_one_launch else 0))      # teacher re-warp + decode + select: one launch
        # the student's grids are built under autocast (:414): every stage samples on a half grid
        stu_half = inp.y_t_stu.dtype in (torch.float16, torch.bfloat16)
        stu_mask = (1 << inp.theta_stu.shape[1]) - 1 if (inp.theta_stu is not None and stu_half) else 0
        stu_grid = inp.y_t_stu.dtype if stu_half else None
        y_t_stu_recon, recon_ready, plan_ready = inp.y_t_stu, None, None
        if inp.theta_stu is not None:
            # the composed map of every sample inverted once — what the backward gathers from.  It depends
            # on theta alone, so it is built on its own branch, off the  gather -> loss -> backward  chain
            # (only with rewarp.USE_INVERSE_PLAN: by default the backward inverts the map itself, in its own launch)
            stu_plan = _rewarp.inverse_plan_buffer(inp.y_t_stu) if _rewarp.USE_INVERSE_PLAN else None
            if stu_plan is not None:
                if self.parallel:
                    s_plan.wait_stream(cur)
                with torch.cuda.stream(s_plan), torch.no_grad():
                    self._mark("plan start")
                    _rewarp.build_inverse_plan(inp.y_t_stu, inp.theta_stu, stu_mask, stu_grid, plan=stu_plan)
                    self._mark("plan done")
                    if self.parallel:
                        plan_ready = torch.cuda.Event()
                        plan_ready.record(s_plan)
            with torch.cuda.stream(s_stu), torch.no_grad():
                # :417-423 — y_t_stu_recon (the backward runs after the loss step, below)
                self._mark("stu gather start")
                y_t_stu_recon = _rewarp.gather(inp.y_t_stu.detach(), inp.theta_stu, stu_mask, stu_grid)
                self._mark("stu gather done")
                if self.parallel:
                    recon_ready = torch.cuda.Event()
                    recon_ready.record(s_stu)
        with torch.cuda.stream(s_tea):
            with torch.no_grad():
                views = inp.teacher_views
                y_t_tea = views[0]
                tt = None
                if tea_one_launch:
                    # :359-383 and :427-430 in one launch; the map itself is never written
                    self._mark("tea gather start")
                    tt = teacher_targets_rewarped(views[0], inp.theta_tea if torch.is_tensor(inp.theta_tea) else inp.theta_tea[0],
                                                  self.sigma, self.mask_ratio, occlude_thresh=self.occlude_thresh)
                    y_t_tea = None
                    self._mark("tea gather done")
                elif inp.theta_tea is not None:
                    # :359-372 — every teacher view warped back to the un-augmented frame, mean over the k views
                    self._mark("tea gather start")
                    if len(views) == 1:
                        y_t_tea = _rewarp.gather(views[0], inp.theta_tea if torch.is_tensor(inp.theta_tea) else inp.theta_tea[0])
                    else:
                        y_t_tea = _rewarp.gather_views(views, list(inp.theta_tea))
                    self._mark("tea gather done")
                elif len(views) > 1:
                    raise ValueError("k teacher views need their re-warp tables (theta_tea: one per view)")
                # train_human.py:376-383 and :427-430 — one decode pass + k-th value select
                if tt is None:
                    tt = teacher_targets(y_t_tea, self.sigma, self.mask_ratio, occlude_thresh=self.occlude_thresh,
                                         materialise=not self.fused)
                self._mark("decode+mask done")
                x_t_stu_occ = None
                if inp.x_t_stu is not None and not torch.cuda.is_current_stream_capturing():
                    # :385-412 — a random patch pasted over one confident keypoint of the student's target image (the
                    # student network runs on the result).  conf_table / position go to the host for the reference's
                    # np.random draws, so this stage exists in eager runs only; a captured step ends before it.
                    x_t_stu_occ = _rewarp.occlude_keypoints(inp.x_t_stu, tt["conf_table"], tt["position"], inp.aug_param_stu,
                                                            self.image_size / views[0].shape[-1], self.occlude_rate,
                                                            self.occlude_size, self.image_size,
                                                            **({"rng": self.rng} if self.rng is not None else {}))
                if recon_ready is not None:
                    s_tea.wait_event(recon_ready)
                if self.fused:
                    # :425-436 — both criteria, loss_all and the scaled-backward seeds in one launch;
                    # the rectified teacher map (:428) is evaluated on the fly from the arg-max
                    losses, g_s, g_c = fused_losses(inp.y_s, inp.label_s, inp.weight_s, y_t_stu_recon, None,
                                                    tt["tea_mask"], lambda_c=self.lambda_c, grad_scale=self.loss_scale,
                                                    tea_preds=tt["preds"], sigma=self.sigma)
                    loss_all, loss_s, loss_c = losses[0], losses[1], losses[2]
                    self._mark("loss step done")
            if not self.fused:
                # :432 — consistency loss; its share of `scaler.scale(loss_all).backward()` (:434-436)
                y_t_stu = y_t_stu_recon.detach().requires_grad_(True)
                loss_c = cons_loss(y_t_stu, tt["rectified"], tea_mask=tt["tea_mask"])
                (g_c,) = torch.autograd.grad(loss_c * (self.lambda_c * self.loss_scale), (y_t_stu,))
            g_recon = g_c
            if inp.theta_stu is not None:
                if plan_ready is not None:
                    s_tea.wait_event(plan_ready)
                with torch.no_grad():
                    # backward of :417-423: the consistency gradient scattered back to the student's frame
                    g_c = _rewarp.gather_backward(g_recon, inp.theta_stu, stu_mask, stu_grid, plan=stu_plan)
                    self._mark("rewarp bwd done")
        with torch.cuda.stream(s_stu):
            if not self.fused:
                # :425 — supervised loss and its share of the scaled backward
                y_s = inp.y_s.detach().requires_grad_(True)
                loss_s = joints_mse_loss(y_s, inp.label_s, inp.weight_s)
                (g_s,) = torch.autograd.grad(loss_s * self.loss_scale, (y_s,))
            with torch.no_grad():
                # :443-444 — PCK on (y_s, label_s): integer counts stay on the device
                self._mark("pck start")
                counts, pred = _pck(inp.y_s, inp.label_s, 0.5)
                self._mark("pck done")
                if self.counts_hook is not None:
                    self.counts_hook(counts)
                pck_done = None
                if ema_side and getattr(self.tail, "exchanges_counts", False):
                    pck_done = torch.cuda.Event()
                    pck_done.record(s_stu)
        t_s2t = t_t2s = None
        if s_adain is not cur:
            s_adain.wait_stream(cur)
        # :348-356 — s2t and t2s feature re-normalisation (the decoder conv follows).  (The two directions on two
        # streams of one priority, so that the second fills the first one's last partial wave: no change, 180.8 vs 180.1 us)
        if "adain" not in self.skip:
            with torch.cuda.stream(s_adain), torch.no_grad():
                if self.adain_one_launch and inp.feat_src.shape == inp.feat_tgt_tea.shape:
                    # both directions are independent: one launch pays the ramp-up / drain of a launch once
                    t_s2t, t_t2s = adain_mix_multi([(inp.feat_src, inp.feat_tgt_ori, inp.alpha_s2t),
                                                    (inp.feat_tgt_tea, inp.feat_src_ori, inp.alpha_t2s)])
                else:
                    t_s2t = adain_mix(inp.feat_src, inp.feat_tgt_ori, inp.alpha_s2t)
                    self._mark("adain s2t done")
                    t_t2s = adain_mix(inp.feat_tgt_tea, inp.feat_src_ori, inp.alpha_t2s)
                self._mark("adain t2s done")
                if self.alpha_feed is not None:
                    self.alpha_feed()
        if ema_side and self.tail is not None:
            with torch.cuda.stream(s_ema):
                # :437-438 — update + EMA; a peer tail first sums the PCK counts over the ranks (same stream: every
                # rank meets its peers in one order)
                if pck_done is not None:
                    s_ema.wait_event(pck_done)
                self._mark("tail start")
                self.tail.finish(counts if pck_done is not None else None)
                self._mark("ema done")
        if s_adain is not cur:
            cur.wait_stream(s_adain)
        if self.parallel:
            cur.wait_stream(s_tea)
            cur.wait_stream(s_stu)
            if ema_side:
                cur.wait_stream(s_ema)
        self._mark("join")
        if not self.fused:
            with torch.no_grad():
                loss_s, loss_c = loss_s.detach(), loss_c.detach()
                loss_all = loss_s + self.lambda_c * loss_c  # :434
        if with_ema and not ema_side:
            self._tail_or_ema()  # :438
        return dict(t_s2t=t_s2t, t_t2s=t_t2s, conf_table=tt["conf_table"], position=tt["position"],
                    tea_mask=tt["tea_mask"], mask_thresh=tt["mask_thresh"], rectified=tt["rectified"],
                    tea_preds=tt["preds"], loss_s=loss_s, loss_c=loss_c, loss_all=loss_all,
                    grad_y_s=g_s, grad_y_t_stu=g_c, grad_y_t_stu_recon=g_recon, y_t_tea_recon=y_t_tea,
                    y_t_stu_recon=y_t_stu_recon, pck_counts=counts, pred=pred, x_t_stu=x_t_stu_occ)

    def _tail_or_ema(self):
        if self.tail is not None:
            if getattr(self.tail, "exchanges_counts", False) and self.tail.world > 1:
                raise RuntimeError("a peer tail needs parallel=True, ema_parallel=True (its count exchange joins the PCK chain)")
            self.tail.run()      # :436-438 — gradient exchange, student update, teacher EMA
        else:
            self.ema.step()      # :438 alone

    def run_no_ema(self, inp: StepInputs) -> dict:
        return self._run(inp, with_ema=False)

    def run(self, inp: StepInputs) -> dict:
        return self._run(inp, with_ema=True)

    # -- CUDA graph of the step (static input/output buffers) ------------------------------------
    def capture(self, inp: StepInputs, include_ema: bool = True, warmup: int = 3) -> dict:
        """Capture the step into a CUDA graph reading from ``inp``'s tensors in place; returns the
        dict of static output tensors that every ``replay()`` overwrites."""
        fn = self.run if include_ema else self.run_no_ema
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(inp)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if self.marks is not None:
            self.marks.clear()     # keep only the events recorded by the captured pass
        with torch.cuda.graph(self.graph):
            self.out = fn(inp)
        self._graph_inputs = inp
        return self.out

    def replay(self) -> dict:
        self.graph.replay()
        return self.out

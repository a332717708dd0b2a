"""Batched nearest-neighbour affine re-warp — the per-sample ``tF.affine`` loops of the trainers
as one gather launch:

* ``train_human.py:361-372`` teacher recon — for every sample and each of the ``k`` teacher
  views three ``tF.affine`` calls (translate/ratio → rotate+scale → shear) through a CPU staging
  tensor with six ``.item()`` syncs, then the mean over views          → :func:`teacher_recon`
* ``train_human.py:418-423`` student recon — the same three calls under autocast, with
  autograd through them                                                  → :func:`student_recon`
* ``train_human.py:385-412`` adaptive occlusion — three-stage warp of the image, patch paste,
  one-stage warp back                                                    → :func:`occlude_keypoints`

(identical code in ``train_animal.py:386-397,443-448,410-437``).  ``tF.affine`` on a tensor is
torchvision's inverse affine matrix (Python float64 ``math``) → ``_gen_affine_grid`` →
``grid_sample(mode="nearest", padding_mode="zeros", align_corners=False)``.  The host part here
builds the per-sample matrices exactly like torchvision (``math`` in float64, rounded to the grid
dtype with the same torch constructors) and ships one tiny ``[B, stages, 6]`` table; the device
composes the stages' integer source-index maps and gathers once (``csrc/rewarp.cu``).

Autocast semantics of the reference are kept (see :func:`stage_table`): inside the trainers'
autocast block every ``bmm`` that builds a sampling grid runs in the autocast dtype, so the student
recon samples on half-precision grids, while the teacher recon and the occlusion (outside the
block) are float32 throughout.
"""
from __future__ import annotations

import ctypes
import os
import math
from typing import Sequence

import numpy as np
import torch

from . import _lib

__all__ = ["inverse_affine_matrix", "recon_stages", "stage_table", "gather", "gather_views", "gather_decode", "gather_decode_supported", "gather_backward", "inverse_plan_buffer", "student_recon", "teacher_recon",
           "affine_nearest", "occlusion_plan", "occlude_keypoints"]


def inverse_affine_matrix(center, angle, translate, scale, shear):
    """torchvision ``transforms.functional._get_inverse_affine_matrix`` (the matrix ``tF.affine``
    hands to ``grid_sample``), float64 ``math`` like torchvision: M^-1 = C · RSS^-1 · C^-1 · T^-1."""
    rot = math.radians(angle)
    sx = math.radians(shear[0])
    sy = math.radians(shear[1])
    cx, cy = center
    tx, ty = translate
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [x / scale for x in m]
    m[2] += m[0] * (-cx - tx) + m[1] * (-cy - ty)
    m[5] += m[3] * (-cx - tx) + m[4] * (-cy - ty)
    m[2] += cx
    m[5] += cy
    return m


def _column(v, n):
    if torch.is_tensor(v):
        v = v.detach().cpu().tolist()
    elif isinstance(v, np.ndarray):
        v = v.tolist()
    if not isinstance(v, (list, tuple)):
        v = [v] * n
    if len(v) != n:
        raise ValueError(f"aug_param column has {len(v)} entries for a batch of {n}")
    return list(v)


def recon_stages(aug_param, ratio: float, batch: int):
    """The three ``tF.affine`` calls of train_human.py:366-368 / :421-423 per sample, in application
    order, as ``(angle, translate, scale, shear)`` tuples.  ``aug_param`` is the collated
    ``meta['aug_param_*']``: ``(angle[B], [trans_x[B], trans_y[B]], [shear_x[B], shear_y[B]], scale[B])``."""
    angle, (trans_x, trans_y), (shear_x, shear_y), scale = aug_param
    cols = [_column(c, batch) for c in (angle, trans_x, trans_y, shear_x, shear_y, scale)]
    out = []
    for ang, tx, ty, sx, sy, sc in zip(*cols):
        out.append([
            (0.0, [tx / ratio, ty / ratio], 1.0, [0.0, 0.0]),
            (float(ang), [0.0, 0.0], sc, [0.0, 0.0]),
            (0.0, [0.0, 0.0], 1.0, [sx, sy]),
        ])
    return out


def _autocast_dtype(autocast, dtype: torch.dtype):
    """``autocast`` argument → the enclosing autocast dtype or None.  "auto" follows an active
    ``torch.autocast('cuda')`` block (the reference's ``torch.cuda.amp.autocast()``)."""
    if autocast == "auto":
        if torch.is_autocast_enabled("cuda"):
            return torch.get_autocast_dtype("cuda")
        if dtype in (torch.float16, torch.bfloat16):
            raise NotImplementedError(
                "re-warp of a half tensor outside autocast is not supported (torchvision would run grid_sample in "
                "half precision; the trainers only do this inside torch.cuda.amp.autocast(), train_human.py:414): "
                "call under autocast or pass autocast=<dtype>")
        return None
    if autocast in (None, False):
        if dtype in (torch.float16, torch.bfloat16):
            raise NotImplementedError("re-warp of a half tensor needs autocast=<dtype> (see rewarp.stage_table)")
        return None
    if autocast not in (torch.float16, torch.bfloat16):
        raise ValueError(f"autocast must be 'auto', None, torch.float16 or torch.bfloat16, got {autocast!r}")
    return autocast


def stage_table(stages, height: int, width: int, dtype: torch.dtype = torch.float32, autocast="auto"):
    """Per-sample stage lists (application order) → ``(theta float32 [B,S,6] in EVALUATION order,
    half_mask, grid_dtype_code)`` for :func:`gather`.

    ``dtype`` is the dtype of the tensor the first stage is applied to.  Outside autocast (teacher
    recon, occlusion) everything is float32.  Inside ``autocast(A)`` (student recon) the first
    ``tF.affine`` builds ``theta`` in ``dtype``; ``grid_sample`` is autocast to float32 and returns
    float32, so the later calls build float32 thetas — and in every call ``bmm`` casts ``theta/(0.5*size)``
    and the base grid to ``A`` and returns the grid in ``A``."""
    n_stage = len(stages[0])
    if any(len(s) != n_stage for s in stages):
        raise ValueError("every sample needs the same number of stages")
    ac = _autocast_dtype(autocast, dtype)
    mats = [[inverse_affine_matrix([0.0, 0.0], ang, [1.0 * t for t in tr], sc, sh) for (ang, tr, sc, sh) in sample]
            for sample in stages]
    table = torch.empty(len(stages), n_stage, 2, 3, dtype=torch.float32)
    for s in range(n_stage):
        td = dtype if (s == 0 or ac is None) else torch.float32
        # torch.tensor(matrix, dtype=img.dtype) and theta.transpose(1, 2) / [0.5*w, 0.5*h]  (F_t.affine,
        # _gen_affine_grid) — the same constructors, so the roundings are torchvision's
        theta = torch.tensor([m[s] for m in mats], dtype=td).reshape(-1, 2, 3)
        denom = torch.tensor([0.5 * width, 0.5 * height], dtype=td)
        r = theta / denom.view(1, 2, 1)
        table[:, s] = (r.to(ac) if ac is not None else r).float()
    table = table.flip(1).reshape(len(stages), n_stage, 6).contiguous()
    half_mask = (1 << n_stage) - 1 if ac is not None else 0
    return table, half_mask, (_lib._DTYPE_CODE[ac] if ac is not None else _lib.F16)


# The backward reads an inverse plan the forward (or build_inverse_plan, on another stream) wrote: per sample a
# shared-memory slot for every output pixel and the source pixels with several contributors, built once per batch.
# UDAPE_REWARP_PLAN=0 makes the backward build and invert the map itself instead (one launch less, but 86-120 us
# against 5.6-25 us at the trainers' / the microbench sizes on B200).
USE_INVERSE_PLAN = os.environ.get("UDAPE_REWARP_PLAN", "1") == "1"


def _launch_fwd(views, thetas, half_mask, grid_code, out, paste=None, paste_after=0, active=None, plan=None):
    y0 = views[0]
    b, c, h, w = y0.shape
    n = len(views)
    in_arr = (ctypes.c_void_p * n)(*[v.data_ptr() for v in views])
    th_arr = (ctypes.c_void_p * n)(*[t.data_ptr() for t in thetas])
    dev = y0.device
    with _lib.on_device(dev):
        st = _lib.load().udape_rewarp_fwd(in_arr, th_arr, n, thetas[0].shape[1], half_mask, grid_code, _lib.ptr(paste),
                                          paste_after, _lib.ptr(active), b, c, h, w, _lib.float_code(y0),
                                          _lib.ptr(out), _lib.ptr(plan), _lib.stream_ptr(dev))
    _lib.check(st, "udape_rewarp_fwd")
    return out


def _check_theta(y, theta):
    if y.dim() != 4:
        raise ValueError(f"rewarp expects [B,C,H,W], got {tuple(y.shape)}")
    if theta.dtype != torch.float32 or theta.dim() != 3 or theta.shape[0] != y.shape[0] or theta.shape[2] != 6:
        raise ValueError(f"theta must be float32 [B={y.shape[0]}, stages, 6], got {theta.dtype} {tuple(theta.shape)}")
    if not (1 <= theta.shape[1] <= 4):
        raise ValueError("1 to 4 stages are supported")


def inverse_plan_buffer(y: torch.Tensor) -> torch.Tensor | None:
    """Device buffer for the inverse plan of a re-warp of ``y`` ([B,C,H,W]) — what the forward saves for
    the backward (a slot per output pixel + the source pixels with several contributors; opaque) — or None
    when the plan route does not apply to this plane size (the backward then inverts the map itself)."""
    b, _, h, w = y.shape
    n = _lib.load().udape_rewarp_plan_elems(h, w, y.element_size())
    return torch.empty((b, n), dtype=torch.int16, device=y.device) if n > 0 else None


def build_inverse_plan(y: torch.Tensor, theta: torch.Tensor, half_mask: int = 0, grid_dtype: torch.dtype | None = None,
                       plan: torch.Tensor | None = None) -> torch.Tensor | None:
    """Fill (and return) the inverse plan of ``gather(y, theta, ...)`` without running the gather: the
    plan depends only on ``theta`` and ``y``'s shape / element size, so it can be built on another stream
    while the forward and the loss run (the backward is its only consumer)."""
    dev = _lib.require_cuda(y, theta, plan)
    _check_theta(y, theta)
    if plan is None:
        plan = inverse_plan_buffer(y)
    if plan is None:
        return None
    grid_code = _lib._DTYPE_CODE[grid_dtype] if grid_dtype is not None else _lib.F16
    _launch_fwd([y], [theta.contiguous()], half_mask, grid_code, None, plan=plan)
    return plan


class _Rewarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, theta, half_mask, grid_code):
        plan = inverse_plan_buffer(y) if USE_INVERSE_PLAN else None
        ctx.save_for_backward(theta, plan)
        ctx.meta = (half_mask, grid_code)
        return _launch_fwd([y], [theta], half_mask, grid_code, torch.empty_like(y), plan=plan)

    @staticmethod
    def backward(ctx, grad_out):
        theta, plan = ctx.saved_tensors
        half_mask, grid_code = ctx.meta
        return _launch_bwd(grad_out.contiguous(), theta, half_mask, grid_code, plan), None, None, None


def _launch_bwd(g, theta, half_mask, grid_code, plan=None):
    b, c, h, w = g.shape
    grad_in = torch.empty_like(g)
    dev = g.device
    with _lib.on_device(dev):
        st = _lib.load().udape_rewarp_bwd(g.data_ptr(), theta.data_ptr(), theta.shape[1], half_mask, grid_code,
                                          b, c, h, w, _lib.float_code(g), grad_in.data_ptr(), _lib.ptr(plan),
                                          _lib.stream_ptr(dev))
    _lib.check(st, "udape_rewarp_bwd")
    return grad_in


def gather_backward(grad_out: torch.Tensor, theta: torch.Tensor, half_mask: int = 0,
                    grid_dtype: torch.dtype | None = None, plan: torch.Tensor | None = None) -> torch.Tensor:
    """Gradient of :func:`gather` w.r.t. its input for an upstream gradient ``grad_out`` (what autograd
    calls; exposed for callers that drive the backward pass themselves, e.g. the fused loss step).
    ``plan``: the inverse plan ``gather(..., plan=...)`` filled for the same ``theta``."""
    _lib.require_cuda(grad_out, theta, plan)
    _check_theta(grad_out, theta)
    grid_code = _lib._DTYPE_CODE[grid_dtype] if grid_dtype is not None else _lib.F16
    return _launch_bwd(grad_out.detach().contiguous(), theta.contiguous(), half_mask, grid_code, plan)


def gather(y: torch.Tensor, theta: torch.Tensor, half_mask: int = 0, grid_dtype: torch.dtype | None = None,
           plan: torch.Tensor | None = None):
    """``out[b,c,p] = y[b,c,source_b(p)]`` for a device stage table from :func:`stage_table`
    (differentiable w.r.t. ``y``; graph-capturable: ``theta`` is a device tensor that can be
    refreshed between replays).  ``plan`` (from :func:`inverse_plan_buffer`): also fill the inverse
    plan for a later :func:`gather_backward` — autograd does this by itself."""
    dev = _lib.require_cuda(y, theta, plan)
    _check_theta(y, theta)
    y = y.contiguous()
    theta = theta.contiguous()
    grid_code = _lib._DTYPE_CODE[grid_dtype] if grid_dtype is not None else _lib.F16
    if y.requires_grad and torch.is_grad_enabled():
        return _Rewarp.apply(y, theta, half_mask, grid_code)
    return _launch_fwd([y], [theta], half_mask, grid_code, torch.empty_like(y), plan=plan)


def gather_decode_supported(y: torch.Tensor) -> bool:
    """Whether :func:`gather_decode` has a fused launch for ``y``: planes of exactly 4096 pixels whose rows are a
    power-of-two number of 16-byte chunks, at most 64 planes per sample (``udape_rewarp_decode_select``)."""
    if y.dim() != 4 or not y.is_cuda or y.dtype not in _lib._DTYPE_CODE:
        return False
    _, c, h, w = y.shape
    row = w * y.element_size()
    cpr = row // 16
    return h * w == 4096 and row % 16 == 0 and cpr >= 1 and (cpr & (cpr - 1)) == 0 and c <= 64 and y.data_ptr() % 16 == 0


def gather_decode(y: torch.Tensor, theta: torch.Tensor, half_mask: int = 0, grid_dtype: torch.dtype | None = None, *,
                  want_idx: bool = False, want_preds: bool = True, want_position: bool = False,
                  occlude_thresh: float | None = None, select_kth: int | None = None,
                  select_tea_mask: torch.Tensor | None = None) -> dict:
    """``decode(gather(y, theta))`` without the re-warped map: one launch (``udape_rewarp_decode_select``) that
    arg-maxes every plane where it is gathered — the teacher chain of the step (train_human.py:359-372 feeding
    :376-383 and :427-430), whose re-warped map is only ever decoded.  Bit-identical to the two calls.  Keys as
    :func:`keypoint_detection.decode`: ``maxvals_f32`` float32[B,K] always, ``idx`` / ``preds`` / ``position`` /
    ``conf_table`` on request, ``tea_mask`` + ``mask_thresh`` with ``select_kth`` (1-based rank, train_human.py:429).
    Raises ``ValueError`` for shapes without a fused launch (see :func:`gather_decode_supported`)."""
    dev = _lib.require_cuda(y, theta, select_tea_mask)
    _check_theta(y, theta)
    _lib.no_autograd("gather_decode", y)
    y = y.detach().contiguous()
    theta = theta.contiguous()
    if not gather_decode_supported(y):
        raise ValueError(f"gather_decode: no fused launch for planes {tuple(y.shape)} {y.dtype}; use gather() then decode()")
    b, k, h, w = y.shape
    planes = b * k
    grid_code = _lib._DTYPE_CODE[grid_dtype] if grid_dtype is not None else _lib.F16
    idx = torch.empty((b, k), dtype=torch.int32, device=dev) if want_idx else None
    preds = torch.empty((b, k, 2), dtype=torch.float32, device=dev) if want_preds else None
    mv32 = torch.empty((b, k), dtype=torch.float32, device=dev)
    pos = torch.empty((b, k, 2), dtype=torch.int64, device=dev) if want_position else None
    conf = torch.empty((b, k), dtype=torch.bool, device=dev) if occlude_thresh is not None else None
    if occlude_thresh is not None and y.dtype != torch.float32:
        occlude_thresh = float(torch.tensor(float(occlude_thresh), dtype=y.dtype))   # torch compares in the tensor's dtype
    kth, tm_in, tea_mask, thresh = 0, None, None, None
    if select_kth is not None:
        if not (1 <= select_kth <= planes):
            raise IndexError(f"kthvalue(): selected number k out of range for dimension 0 (k={select_kth}, n={planes})")
        kth = int(select_kth)
        if select_tea_mask is not None:
            if select_tea_mask.numel() != planes:
                raise ValueError("gather_decode: select_tea_mask must have one entry per (b, k)")
            tm_in = select_tea_mask.detach().to(torch.float32).contiguous()
        tea_mask = torch.empty((b, k), dtype=torch.bool, device=dev)
        thresh = torch.empty((), dtype=torch.float32, device=dev)
    with _lib.on_device(dev):
        st = _lib.load().udape_rewarp_decode_select(
            y.data_ptr(), theta.data_ptr(), theta.shape[1], int(half_mask), grid_code, b, k, h, w, _lib.float_code(y),
            _lib.ptr(idx), _lib.ptr(preds), mv32.data_ptr(), _lib.ptr(pos),
            float(occlude_thresh if occlude_thresh is not None else 0.0), _lib.ptr(conf), kth, _lib.ptr(tm_in),
            _lib.ptr(thresh), _lib.ptr(tea_mask), _lib.ticket(dev) if kth else None, _lib.stream_ptr(dev))
    _lib.check(st, "gather_decode")
    out = {"maxvals_f32": mv32}
    for name, t in (("idx", idx), ("preds", preds), ("position", pos), ("conf_table", conf), ("tea_mask", tea_mask),
                    ("mask_thresh", thresh)):
        if t is not None:
            out[name] = t
    return out


def gather_views(views: Sequence[torch.Tensor], thetas: Sequence[torch.Tensor], half_mask: int = 0,
                 grid_dtype: torch.dtype | None = None) -> torch.Tensor:
    """Mean over ``k`` views of ``gather(view_i, theta_i)`` in one launch (``teacher_recon`` with device stage
    tables: graph-capturable, the tables can be refreshed between replays).  Forward only."""
    if len(views) != len(thetas) or not (1 <= len(views) <= 4):
        raise ValueError("gather_views: one stage table per view, 1 to 4 views")
    _lib.require_cuda(*views, *thetas)
    _lib.no_autograd("gather_views", *views)
    vs = [v.detach().contiguous() for v in views]
    if any(v.shape != vs[0].shape or v.dtype != vs[0].dtype for v in vs):
        raise ValueError("gather_views: all views must share shape and dtype")
    for v, t in zip(vs, thetas):
        _check_theta(v, t)
    grid_code = _lib._DTYPE_CODE[grid_dtype] if grid_dtype is not None else _lib.F16
    return _launch_fwd(vs, [t.contiguous() for t in thetas], half_mask, grid_code, torch.empty_like(vs[0]))


def student_recon(y_t_stu: torch.Tensor, aug_param_stu, ratio: float, autocast="auto") -> torch.Tensor:
    """``y_t_stu_recon`` of train_human.py:418-423: the student's target heatmaps warped back with
    the sample's augmentation parameters; autograd flows to ``y_t_stu``.  Call it where the reference
    runs the loop — inside the autocast block — or pass ``autocast=torch.float16`` explicitly."""
    _lib.require_cuda(y_t_stu)
    b, _, h, w = y_t_stu.shape
    table, half_mask, code = stage_table(recon_stages(aug_param_stu, ratio, b), h, w, y_t_stu.dtype, autocast)
    theta = table.to(y_t_stu.device, non_blocking=True)
    yc = y_t_stu.contiguous()
    if half_mask and yc.dtype == torch.float32:
        # torchvision casts a float32 image to the (half) grid dtype before sampling and back after
        yc = yc.to(_autocast_dtype(autocast, yc.dtype)).float()
    if yc.requires_grad and torch.is_grad_enabled():
        return _Rewarp.apply(yc, theta, half_mask, code)
    return _launch_fwd([yc], [theta], half_mask, code, torch.empty_like(yc))


def teacher_recon(y_t_teas: Sequence[torch.Tensor], aug_params_tea: Sequence, ratio: float) -> torch.Tensor:
    """``y_t_tea_recon`` of train_human.py:359-372: every teacher view warped back and the mean over
    the ``k`` views (forward only — the trainers run it under ``torch.no_grad()``)."""
    if len(y_t_teas) != len(aug_params_tea) or not (1 <= len(y_t_teas) <= 4):
        raise ValueError("teacher_recon: one aug_param per view, 1 to 4 views")
    dev = _lib.require_cuda(*y_t_teas)
    _lib.no_autograd("teacher_recon", *y_t_teas)
    b, _, h, w = y_t_teas[0].shape
    views = [v.detach().contiguous() for v in y_t_teas]
    if any(v.shape != views[0].shape or v.dtype != views[0].dtype for v in views):
        raise ValueError("teacher_recon: all views must share shape and dtype")
    thetas, masks = [], set()
    for ap in aug_params_tea:
        table, half_mask, code = stage_table(recon_stages(ap, ratio, b), h, w, views[0].dtype, None)
        thetas.append(table.to(dev, non_blocking=True))
        masks.add((half_mask, code))
    half_mask, code = masks.pop()
    return _launch_fwd(views, thetas, half_mask, code, torch.empty_like(views[0]))


def affine_nearest(img: torch.Tensor, angle, translate, scale, shear, autocast="auto") -> torch.Tensor:
    """Batched ``tF.affine(img, angle, translate, scale, shear)`` (nearest, zero fill) for a
    ``[B,C,H,W]`` tensor with per-sample parameters (scalars broadcast), or one ``[C,H,W]`` image."""
    squeeze = img.dim() == 3
    x = img.unsqueeze(0) if squeeze else img
    _lib.require_cuda(x)
    _lib.no_autograd("affine_nearest", x)
    b, _, h, w = x.shape
    ang = _column(angle, b)
    sc = _column(scale, b)
    tr = translate if isinstance(translate[0], (list, tuple)) else [translate] * b
    sh = shear if isinstance(shear[0], (list, tuple)) else [shear] * b
    stages = [[(float(ang[i]), list(tr[i]), sc[i], list(sh[i]))] for i in range(b)]
    table, half_mask, code = stage_table(stages, h, w, x.dtype, autocast)
    out = _launch_fwd([x.contiguous()], [table.to(x.device, non_blocking=True)], half_mask, code, torch.empty_like(x))
    return out.squeeze(0) if squeeze else out


def occlusion_plan(conf_table, pred_position, aug_param_stu, ratio: float, occlude_rate: float, occlude_size: int,
                   image_size: int, rng=np.random):
    """Host part of train_human.py:385-412: draws, in the reference's order and from the same
    ``np.random`` stream, which samples are occluded, the occluded joint and the source patch.
    ``conf_table`` bool [B,K] and ``pred_position`` int [B,K,2] (x, y) are host arrays (the reference
    moves both to the host too).  Returns ``(active uint8 [B], paste int32 [B,6], stages)`` where
    ``stages[b]`` lists the four tF.affine calls (three-stage warp, then the warp back)."""
    conf_table = np.asarray(conf_table).astype(bool)
    pred_position = np.asarray(pred_position)
    b, k = conf_table.shape
    angle, (trans_x, trans_y), (shear_x, shear_y), scale = aug_param_stu
    cols = [_column(c, b) for c in (angle, trans_x, trans_y, shear_x, shear_y, scale)]
    active = np.zeros(b, dtype=np.uint8)
    paste = np.zeros((b, 6), dtype=np.int32)
    stages = []
    for _b in range(b):
        ang, tx, ty, sx, sy, sc = (c[_b] for c in cols)
        stages.append([
            (0.0, [tx / ratio, ty / ratio], 1.0, [0.0, 0.0]),
            (float(ang), [0.0, 0.0], sc, [0.0, 0.0]),
            (0.0, [0.0, 0.0], 1.0, [sx, sy]),
            (float(-ang), [-tx / ratio, -ty / ratio], 1.0 / sc, [-sx, -sy]),  # :412 warp it back
        ])
        if conf_table[_b].sum() > 0 and rng.rand() <= occlude_rate:
            candidates = np.arange(0, k)[conf_table[_b]]
            _c = rng.choice(candidates)
            position = (pred_position[_b, _c] * ratio).astype(int)
            left = max(position[1] - occlude_size, 0)
            right = min(position[1] + occlude_size, image_size)
            upper = max(position[0] - occlude_size, 0)
            bottom = min(position[0] + occlude_size, image_size)
            left_src = rng.randint(image_size - (right - left) + 1)
            upper_src = rng.randint(image_size - (bottom - upper) + 1)
            active[_b] = 1
            paste[_b] = (left, right, upper, bottom, left_src, upper_src)
    return active, paste, stages


def occlude_keypoints(x_t_stu: torch.Tensor, conf_table, pred_position, aug_param_stu, ratio: float,
                      occlude_rate: float, occlude_size: int, image_size: int, rng=np.random) -> torch.Tensor:
    """train_human.py:385-412 for the whole batch in one launch: selected samples are warped to the
    teacher frame, a random patch is pasted over the chosen keypoint and the image is warped back;
    the others are returned unchanged.  Returns a new tensor (the reference assigns in place).

    The reference's patch copy raises for overlapping source/destination patches (torch refuses
    partially overlapping ``copy_``); this operator reads the pre-paste image instead."""
    _lib.require_cuda(x_t_stu)
    _lib.no_autograd("occlude_keypoints", x_t_stu)
    if torch.is_tensor(conf_table):
        conf_table = conf_table.detach().cpu().numpy()
    if torch.is_tensor(pred_position):
        pred_position = pred_position.detach().cpu().numpy()
    active, paste, stages = occlusion_plan(conf_table, pred_position, aug_param_stu, ratio, occlude_rate,
                                           occlude_size, image_size, rng)
    x = x_t_stu.detach().contiguous()
    if not active.any():
        return x.clone()
    b, _, h, w = x.shape
    table, half_mask, code = stage_table(stages, h, w, x.dtype, None)  # :385 runs outside the autocast block
    dev = x.device
    # evaluation order: warp-back first, then the paste remap, then shear, rotate+scale, translate
    return _launch_fwd([x], [table.to(dev, non_blocking=True)], half_mask, code, torch.empty_like(x),
                       paste=torch.from_numpy(paste).to(dev, non_blocking=True), paste_after=1,
                       active=torch.from_numpy(active).to(dev, non_blocking=True))

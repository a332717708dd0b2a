"""Heatmap losses — drop-in for ``lib/models/loss.py`` of the reference:
``JointsMSELoss`` (:11-49) and ``ConsLoss`` (:119-132).

Forward and backward are hand-written sm_100a kernels (``csrc/loss.cu``) behind
``torch.autograd.Function``s, valid under ``torch.cuda.amp.autocast`` + ``GradScaler``:
the student heatmap may be fp16/bf16 while the label / rectified teacher map is fp32; sums
are accumulated in fp32, the loss is fp32, and the gradient is returned in the student
heatmap's dtype (what autograd would hand back through autocast's cast).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

__all__ = ["JointsMSELoss", "ConsLoss", "joints_mse_loss", "cons_loss", "fused_losses"]


def _check_pair(name, a, b):
    if a.dim() != 4:
        raise ValueError(f"{name}: expected [B,K,H,W] heatmaps, got {tuple(a.shape)}")
    if a.shape != b.shape:
        raise ValueError(f"{name}: shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    return _lib.require_cuda(a, b)


def _no_grad_operand(name, what, t):
    if t is not None and t.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError(
            f"{name}: gradients w.r.t. `{what}` are not implemented (the trainers never need them: "
            "labels and teacher heatmaps are constants, train_human.py:347-358,425-432); detach it")


def _result_dtype(a, b):
    """The reference's loss dtype: fp32 under autocast (mse_loss / pow are fp32 ops), else the
    promoted dtype of the operands."""
    if torch.is_autocast_enabled():
        return torch.float32
    return torch.promote_types(a.dtype, b.dtype)


class _JointsMSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, weight, per_plane: bool):
        dev = output.device
        b, k, h, w = output.shape
        planes, hw = b * k, h * w
        output = output.contiguous()
        target = target.contiguous()
        if weight is not None:
            weight = weight.contiguous()
        # one scratch allocation: plane_loss[planes] | loss
        scratch = torch.empty(planes + 1, dtype=torch.float32, device=dev)
        w_code = _lib.dtype_code(weight) if weight is not None else 0
        base = scratch.data_ptr()
        with _lib.on_device(dev):
            st = _lib.load().udape_joints_mse_fwd(
                output.data_ptr(), _lib.float_code(output), target.data_ptr(), _lib.float_code(target),
                _lib.ptr(weight), w_code, planes, hw, base,
                None if per_plane else base + 4 * planes, None if per_plane else _lib.ticket(dev),
                _lib.stream_ptr(dev))
        _lib.check(st, "JointsMSELoss.forward")
        ctx.save_for_backward(output, target, weight)
        ctx.per_plane = per_plane
        return scratch[:planes].view(b, k) if per_plane else scratch[planes]

    @staticmethod
    def backward(ctx, grad_out):
        output, target, weight = ctx.saved_tensors
        dev = output.device
        b, k, h, w = output.shape
        g = grad_out.detach().to(dtype=torch.float32).contiguous()
        grad_in = torch.empty_like(output)
        w_code = _lib.dtype_code(weight) if weight is not None else 0
        with _lib.on_device(dev):
            st = _lib.load().udape_joints_mse_bwd(
                output.data_ptr(), _lib.float_code(output), target.data_ptr(), _lib.float_code(target),
                _lib.ptr(weight), w_code, b * k, h * w, g.data_ptr(), 1 if ctx.per_plane else 0,
                grad_in.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(st, "JointsMSELoss.backward")
        return grad_in, None, None, None


def joints_mse_loss(output, target, target_weight=None, reduction="mean"):
    """Functional form of :class:`JointsMSELoss`."""
    if reduction not in ("mean", "none"):
        raise ValueError(f"JointsMSELoss: unknown reduction {reduction!r}")
    dev = _check_pair("JointsMSELoss", output, target)
    _no_grad_operand("JointsMSELoss", "target", target)
    _no_grad_operand("JointsMSELoss", "target_weight", target_weight)
    b, k = output.shape[:2]
    weight = None
    if target_weight is not None:
        _lib.require_cuda(output, target_weight)
        if target_weight.numel() != b * k:  # loss.py:45 views it (B, K, 1)
            raise RuntimeError(f"JointsMSELoss: target_weight has {target_weight.numel()} elements, expected B*K = {b * k}")
        weight = target_weight.detach().reshape(b * k)
        if weight.dtype not in (torch.float32, torch.float16, torch.bfloat16, torch.bool, torch.uint8):
            weight = weight.float()
    if output.numel() == 0:
        raise ValueError("JointsMSELoss: empty input")
    loss = _JointsMSEFn.apply(output, target.detach(), weight, reduction == "none")
    rd = _result_dtype(output, target)
    if weight is not None and weight.dtype.is_floating_point and not torch.is_autocast_enabled():
        rd = torch.promote_types(rd, weight.dtype)
    return loss if rd == torch.float32 else loss.to(rd)


class JointsMSELoss(nn.Module):
    """``0.5 * (output - target)^2 * target_weight[b,k]`` averaged over everything
    (``reduction='mean'``, scalar) or over each plane (``'none'``, ``[B,K]``) — loss.py:11-49."""

    def __init__(self, reduction="mean"):
        super().__init__()
        # kept for the module's attribute / repr / state_dict contract (loss.py:36); forward() does not call it
        self.criterion = nn.MSELoss(reduction="none")
        self.reduction = reduction

    def forward(self, output, target, target_weight=None):
        return joints_mse_loss(output, target, target_weight, self.reduction)


class _ConsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stu, tea, tea_mask, valid_mask):
        dev = stu.device
        b, k, h, w = stu.shape
        planes, hw = b * k, h * w
        stu = stu.contiguous()
        tea = tea.contiguous()
        # scratch: plane_partial[planes] | loss | (unused) | valid_count
        scratch = torch.empty(planes + 3, dtype=torch.float32, device=dev)
        m_code = _lib.dtype_code(tea_mask) if tea_mask is not None else 0
        base = scratch.data_ptr()
        with _lib.on_device(dev):
            st = _lib.load().udape_cons_fwd(
                stu.data_ptr(), _lib.float_code(stu), tea.data_ptr(), _lib.float_code(tea),
                _lib.ptr(tea_mask), m_code, _lib.ptr(valid_mask), b, k, hw, base,
                base + 4 * (planes + 2), base + 4 * planes, _lib.ticket(dev),
                _lib.stream_ptr(dev))
        _lib.check(st, "ConsLoss.forward")
        ctx.save_for_backward(stu, tea, tea_mask, valid_mask, scratch)
        return scratch[planes]

    @staticmethod
    def backward(ctx, grad_out):
        stu, tea, tea_mask, valid_mask, scratch = ctx.saved_tensors
        dev = stu.device
        b, k, h, w = stu.shape
        planes = b * k
        g = grad_out.detach().to(dtype=torch.float32).contiguous()
        grad_stu = torch.empty_like(stu)
        m_code = _lib.dtype_code(tea_mask) if tea_mask is not None else 0
        with _lib.on_device(dev):
            st = _lib.load().udape_cons_bwd(
                stu.data_ptr(), _lib.float_code(stu), tea.data_ptr(), _lib.float_code(tea),
                _lib.ptr(tea_mask), m_code, _lib.ptr(valid_mask), b, k, h * w, g.data_ptr(),
                scratch.data_ptr() + 4 * (planes + 2), grad_stu.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(st, "ConsLoss.backward")
        return grad_stu, None, None, None


def cons_loss(stu_out, tea_out, valid_mask=None, tea_mask=None):
    """Functional form of :class:`ConsLoss`."""
    dev = _check_pair("ConsLoss", stu_out, tea_out)
    _no_grad_operand("ConsLoss", "tea_out", tea_out)
    b, k, h, w = stu_out.shape
    if stu_out.numel() == 0:
        raise ValueError("ConsLoss: empty input")
    tm = None
    if tea_mask is not None:
        _lib.require_cuda(stu_out, tea_mask)
        if tuple(tea_mask.shape) != (b, k):  # loss.py:127 indexes it [:, :, None, None]
            raise RuntimeError(f"ConsLoss: tea_mask must be [B,K] = {(b, k)}, got {tuple(tea_mask.shape)}")
        tm = tea_mask.detach().contiguous()
        if tm.dtype not in (torch.bool, torch.uint8, torch.float32):
            tm = tm.float()
    vm = None
    if valid_mask is not None:
        _lib.require_cuda(stu_out, valid_mask)
        if valid_mask.dtype != torch.bool or tuple(valid_mask.shape) != (b, h, w):  # loss.py:129-130
            raise RuntimeError(f"ConsLoss: valid_mask must be a bool [B,H,W] = {(b, h, w)} tensor")
        vm = valid_mask.contiguous()
    loss = _ConsFn.apply(stu_out, tea_out.detach(), tm, vm)
    rd = _result_dtype(stu_out, tea_out)
    return loss if rd == torch.float32 else loss.to(rd)


class ConsLoss(nn.Module):
    """Teacher–student consistency: ``mean(((stu - tea) * tea_mask[b,k])^2)`` over k then over
    the (optionally ``valid_mask``-selected) positions — loss.py:119-132."""

    def __init__(self):
        super().__init__()

    def forward(self, stu_out, tea_out, valid_mask=None, tea_mask=None):
        return cons_loss(stu_out, tea_out, valid_mask=valid_mask, tea_mask=tea_mask)


def fused_losses(y_s, label_s, weight_s, y_t_stu, tea_out=None, tea_mask=None, lambda_c: float = 1.0,
                 grad_scale=1.0, tea_preds=None, sigma=None, want_grads: bool = True):
    """Both criteria of the mean-teacher step and the seed of its scaled backward in ONE launch
    (``train_human.py:425-436``)::

        loss_s   = JointsMSELoss()(y_s, label_s, weight_s)
        loss_c   = ConsLoss()(y_t_stu, tea_out, tea_mask=tea_mask)
        loss_all = loss_s + lambda_c * loss_c
        scaler.scale(loss_all).backward()        # -> d(scale*loss_all)/d y_s, /d y_t_stu

    Returns ``(losses, grad_y_s, grad_y_t_stu)``: ``losses`` is a float32 ``[3]`` device tensor
    ``(loss_all, loss_s, loss_c)``; the gradients have the student heatmaps' dtype and are what
    ``torch.autograd.backward([y_s, y_t_stu], [grad_y_s, grad_y_t_stu])`` needs to continue into the
    student network.  ``grad_scale`` is a float or a 1-element float32 CUDA tensor (GradScaler's
    scale).  Instead of a materialised ``tea_out = rectify(y_t_tea, sigma)`` the caller may pass
    ``tea_preds`` (``decode(...)["preds"]`` of the teacher heatmaps) and ``sigma``: the rectified
    map is then evaluated on the fly, never written to or read from HBM, with identical values.
    """
    dev = _lib.require_cuda(y_s, label_s, weight_s, y_t_stu, tea_out, tea_mask, tea_preds)
    have_s, have_t = y_s is not None, y_t_stu is not None
    if not (have_s or have_t):
        raise ValueError("fused_losses: at least one of (y_s, label_s) / (y_t_stu, tea) is required")
    stu = y_s if have_s else y_t_stu
    b_s = k_s = b_t = 0
    if have_s:
        _check_pair("fused_losses", y_s, label_s)
        b_s, k_s, h, w = y_s.shape
    if have_t:
        if y_t_stu.dim() != 4:
            raise ValueError(f"fused_losses: expected [B,K,H,W] heatmaps, got {tuple(y_t_stu.shape)}")
        b_t, k, h2, w2 = y_t_stu.shape
        if have_s and (h2, w2) != (h, w):
            raise ValueError("fused_losses: y_s and y_t_stu must share the heatmap size")
        h, w = h2, w2
        if y_t_stu.dtype != stu.dtype:
            raise TypeError("fused_losses: y_s and y_t_stu must share a dtype")
        if tea_out is None:
            if tea_preds is None or sigma is None:
                raise ValueError("fused_losses: pass either tea_out or (tea_preds, sigma)")
            if tuple(tea_preds.shape) != (b_t, k, 2) or tea_preds.dtype != torch.float32:
                raise ValueError("fused_losses: tea_preds must be float32 [B,K,2]")
            tea_preds = tea_preds.detach().contiguous()
        else:
            if tea_out.shape != y_t_stu.shape:
                raise ValueError(f"fused_losses: shape mismatch {tuple(y_t_stu.shape)} vs {tuple(tea_out.shape)}")
            tea_out = tea_out.detach().contiguous()
            if have_s and tea_out.dtype != label_s.dtype:
                raise TypeError("fused_losses: label_s and tea_out must share a dtype")
    else:
        k = k_s
    tgt_dtype = label_s.dtype if have_s else (tea_out.dtype if tea_out is not None else torch.float32)
    weight = None
    if have_s and weight_s is not None:
        if weight_s.numel() != b_s * k_s:
            raise RuntimeError(f"fused_losses: target_weight has {weight_s.numel()} elements, expected {b_s * k_s}")
        weight = weight_s.detach().reshape(b_s * k_s).contiguous()
        if weight.dtype not in (torch.float32, torch.float16, torch.bfloat16, torch.bool, torch.uint8):
            weight = weight.float()
    tm = None
    if have_t and tea_mask is not None:
        if tuple(tea_mask.shape) != (b_t, k):
            raise RuntimeError(f"fused_losses: tea_mask must be [B,K] = {(b_t, k)}, got {tuple(tea_mask.shape)}")
        tm = tea_mask.detach().contiguous()
        if tm.dtype not in (torch.bool, torch.uint8, torch.float32):
            tm = tm.float()
    gs_dev, gs_host = None, 1.0
    if isinstance(grad_scale, torch.Tensor):
        if not (grad_scale.is_cuda and grad_scale.dtype == torch.float32 and grad_scale.numel() == 1):
            raise TypeError("fused_losses: a tensor grad_scale must be a 1-element float32 CUDA tensor")
        gs_dev = grad_scale.data_ptr()
    else:
        gs_host = float(grad_scale)
    planes_s, planes_t = b_s * k_s, b_t * k
    y_s_c = y_s.detach().contiguous() if have_s else None
    y_t_c = y_t_stu.detach().contiguous() if have_t else None
    label_c = label_s.detach().contiguous() if have_s else None
    # scratch: partial[planes_s + planes_t] | losses[3]
    scratch = torch.empty(planes_s + planes_t + 3, dtype=torch.float32, device=dev)
    base = scratch.data_ptr()
    g_s = torch.empty_like(y_s_c) if (have_s and want_grads) else None
    g_t = torch.empty_like(y_t_c) if (have_t and want_grads) else None
    with _lib.on_device(dev):
        st = _lib.load().udape_loss_step(
            _lib.ptr(y_s_c), _lib.ptr(label_c), _lib.ptr(weight), _lib.dtype_code(weight) if weight is not None else 0,
            planes_s, _lib.ptr(y_t_c), _lib.ptr(tea_out), _lib.ptr(tea_preds), float(sigma) if sigma is not None else 1.0,
            _lib.ptr(tm), _lib.dtype_code(tm) if tm is not None else 0, b_t, k, h, w, _lib.float_code(stu),
            _lib._DTYPE_CODE[tgt_dtype], float(lambda_c), gs_host, gs_dev, base, base + 4 * (planes_s + planes_t),
            _lib.ticket(dev), _lib.ptr(g_s), _lib.ptr(g_t), _lib.stream_ptr(dev))
    _lib.check(st, "fused_losses")
    return scratch[planes_s + planes_t:planes_s + planes_t + 3], g_s, g_t

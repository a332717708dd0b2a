"""AdaIN channel statistics and re-normalisation — drop-in for the reference's
``adain/function.py`` (``calc_mean_std`` :3-11, ``adaptive_instance_normalization`` :14-22)
and the identical copies in ``lib/models/Style_net.py`` (:4-12, ``adain`` :21-29) plus the
s2t/t2s alpha mixing line ``t = alpha * t + (1 - alpha) * content_feat`` (:167-168).

Each call is ONE hand-written sm_100a kernel launch (``csrc/adain.cu``) instead of the
reference's ~19 eager passes.  CUDA tensors only; fp32 / fp16 / bf16.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["calc_mean_std", "calc_style_loss", "adaptive_instance_normalization", "adain", "adain_mix", "adain_mix_multi", "channel_clamp"]


def _planes(feat: torch.Tensor, name: str):
    if feat.dim() != 4:  # function.py:6 `assert (len(size) == 4)`
        raise AssertionError(f"{name}: expected a 4-D NCHW tensor, got {tuple(feat.shape)}")
    n, c, h, w = feat.shape
    return n, c, h * w


def _mean_std_launch(feat: torch.Tensor, eps: float):
    n, c, h, w = feat.shape
    dev = feat.device
    out = torch.empty((2, n, c, 1, 1), dtype=feat.dtype, device=dev)
    if n * c > 0:
        with _lib.on_device(dev):
            st = _lib.load().udape_mean_std(feat.data_ptr(), _lib.float_code(feat), n * c, h * w, float(eps),
                                            out[0].data_ptr(), out[1].data_ptr(), _lib.stream_ptr(dev))
        _lib.check(st, "calc_mean_std")
    return out[0], out[1]


class _MeanStd(torch.autograd.Function):
    """calc_mean_std with a hand-written backward (``udape_mean_std_bwd``): what the style loss of the
    AdaIN decoder pre-training job differentiates through (adain/net.py:137-143)."""

    @staticmethod
    def forward(ctx, feat, eps):
        mean, std = _mean_std_launch(feat, eps)
        ctx.save_for_backward(feat, mean, std)
        return mean, std

    @staticmethod
    def backward(ctx, dmean, dstd):
        feat, mean, std = ctx.saved_tensors
        n, c, h, w = feat.shape
        dev = feat.device
        dfeat = torch.empty_like(feat)
        if feat.numel() > 0:
            dm = dmean.to(feat.dtype).contiguous() if dmean is not None else None
            ds = dstd.to(feat.dtype).contiguous() if dstd is not None else None
            with _lib.on_device(dev):
                st = _lib.load().udape_mean_std_bwd(feat.data_ptr(), mean.data_ptr(), std.data_ptr(), _lib.ptr(dm),
                                                    _lib.ptr(ds), _lib.float_code(feat), n * c, h * w,
                                                    dfeat.data_ptr(), _lib.stream_ptr(dev))
            _lib.check(st, "calc_mean_std (backward)")
        return dfeat, None


def calc_mean_std(feat: torch.Tensor, eps: float = 1e-5):
    """Per-(n,c) mean and ``sqrt(unbiased var + eps)`` over H*W → two ``[N,C,1,1]`` tensors.

    Same signature and return convention as ``adain/function.py:3-11``.  Differentiable: when ``feat``
    requires grad the backward is one elementwise launch (``adain/net.py:137-143`` back-propagates
    the style loss through these statistics into the decoder).
    """
    _planes(feat, "calc_mean_std")
    _lib.require_cuda(feat)
    _lib.float_code(feat)
    feat = feat.contiguous()
    if feat.requires_grad and torch.is_grad_enabled():
        return _MeanStd.apply(feat, float(eps))
    return _mean_std_launch(feat, eps)


def calc_style_loss(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """``Net.calc_style_loss`` of the decoder pre-training job (adain/net.py:137-143):
    ``mse(mean(input), mean(target)) + mse(std(input), std(target))``; gradients flow to ``input``
    only (the reference asserts ``target.requires_grad is False``)."""
    assert input.size() == target.size()
    assert target.requires_grad is False
    input_mean, input_std = calc_mean_std(input)
    target_mean, target_std = calc_mean_std(target)
    return torch.nn.functional.mse_loss(input_mean, target_mean) + torch.nn.functional.mse_loss(input_std, target_std)


def adain_mix(content_feat: torch.Tensor, style_feat: torch.Tensor, alpha=1.0, eps: float = 1e-5,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """Fused ``alpha * adain(content, style) + (1 - alpha) * content`` (Style_net.py:167-168).

    ``alpha`` is a Python float in [0, 1] (asserted like Style_net.py:164) or a 0-dim / 1-element
    float32 CUDA tensor, in which case the kernel reads it from device memory so that a
    captured CUDA graph can be replayed with a new alpha every step.
    """
    if content_feat.shape[:2] != style_feat.shape[:2]:  # function.py:15
        raise AssertionError(
            f"adain: content {tuple(content_feat.shape)} and style {tuple(style_feat.shape)} "
            "must agree in N and C")
    n, c, hw_c = _planes(content_feat, "adain")
    _, _, hw_s = _planes(style_feat, "adain")
    dev = _lib.require_cuda(content_feat, style_feat)
    _lib.no_autograd("adain", content_feat, style_feat)
    if content_feat.dtype != style_feat.dtype:
        raise TypeError(f"adain: dtype mismatch {content_feat.dtype} vs {style_feat.dtype}")
    content_feat = content_feat.contiguous()
    style_feat = style_feat.contiguous()
    code = _lib.float_code(content_feat)
    alpha_dev = None
    alpha_host = 1.0
    if isinstance(alpha, torch.Tensor):
        if not (alpha.is_cuda and alpha.dtype == torch.float32 and alpha.numel() == 1):
            raise TypeError("adain_mix: a tensor alpha must be a 1-element float32 CUDA tensor")
        alpha_dev = alpha.data_ptr()
    else:
        alpha_host = float(alpha)
        assert 0 <= alpha_host <= 1  # Style_net.py:164
    if out is None:
        out = torch.empty_like(content_feat)
    elif out.shape != content_feat.shape or out.dtype != content_feat.dtype or not out.is_contiguous():
        raise ValueError("adain_mix: `out` must be a contiguous tensor shaped and typed like content_feat")
    if n * c > 0:
        with _lib.on_device(dev):
            st = _lib.load().udape_adain_mix(content_feat.data_ptr(), style_feat.data_ptr(), code, n * c,
                                             hw_c, hw_s, float(eps), alpha_host, alpha_dev,
                                             out.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(st, "adain")
    return out


def adain_mix_multi(jobs, eps: float = 1e-5):
    """Several independent ``adain_mix`` calls of ONE shape and dtype as one launch — the s2t and the t2s direction
    of a train step (``train_human.py:348-356``).  ``jobs``: sequence of ``(content, style, alpha)`` (alpha: float or
    1-element float32 CUDA tensor, as in :func:`adain_mix`); returns the list of outputs."""
    import ctypes

    jobs = list(jobs)
    if not 1 <= len(jobs) <= 4:
        raise ValueError("adain_mix_multi: 1..4 jobs")
    c0, s0, _ = jobs[0]
    n, c, hw_c = _planes(c0, "adain")
    _, _, hw_s = _planes(s0, "adain")
    dev = _lib.require_cuda(*[t for j in jobs for t in j[:2]])
    code = _lib.float_code(c0)
    table = (_lib.AdainJob * len(jobs))()
    outs, keep = [], []
    for i, (content, style, alpha) in enumerate(jobs):
        if content.shape != c0.shape or style.shape != s0.shape or content.dtype != c0.dtype or style.dtype != c0.dtype:
            raise ValueError("adain_mix_multi: every job must have the shapes and dtype of the first")
        if content.shape[:2] != style.shape[:2]:  # function.py:15
            raise AssertionError("adain: content and style must agree in N and C")
        _lib.no_autograd("adain", content, style)
        content, style = content.contiguous(), style.contiguous()
        out = torch.empty_like(content)
        keep += [content, style]
        table[i].content, table[i].style, table[i].out = content.data_ptr(), style.data_ptr(), out.data_ptr()
        if isinstance(alpha, torch.Tensor):
            if not (alpha.is_cuda and alpha.dtype == torch.float32 and alpha.numel() == 1):
                raise TypeError("adain_mix: a tensor alpha must be a 1-element float32 CUDA tensor")
            table[i].alpha_dev, table[i].alpha = alpha.data_ptr(), 1.0
        else:
            assert 0 <= float(alpha) <= 1  # Style_net.py:164
            table[i].alpha_dev, table[i].alpha = None, float(alpha)
        outs.append(out)
    if n * c > 0:
        with _lib.on_device(dev):
            st = _lib.load().udape_adain_mix_multi(table, len(jobs), code, n * c, hw_c, hw_s, float(eps), _lib.stream_ptr(dev))
        _lib.check(st, "adain")
    return outs


def adaptive_instance_normalization(content_feat: torch.Tensor, style_feat: torch.Tensor) -> torch.Tensor:
    """``(content - mean_c) / std_c * std_s + mean_s`` per (n,c) plane (function.py:14-22)."""
    return adain_mix(content_feat, style_feat, 1.0)


# lib/models/Style_net.py:21 names the same function `adain`
adain = adaptive_instance_normalization


def channel_clamp(x: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Per-channel clamp of a stylised image batch, the expression every trainer applies to the
    style-transfer output (``train_human.py:276,351,356``)::

        torch.maximum(torch.minimum(x.permute(0,2,3,1), recover_max), recover_min).permute(0,3,1,2)

    ``x`` is ``[N,C,H,W]``; ``lo`` / ``hi`` are the ``[C]`` tensors ``recover_min`` / ``recover_max``.
    One contiguous pass (the reference makes two permuted ones and returns a channels-last view;
    this returns a contiguous NCHW tensor with the same values).  ``out=x`` clamps in place.
    """
    if x.dim() != 4:
        raise ValueError(f"channel_clamp: expected [N,C,H,W], got {tuple(x.shape)}")
    n, c, h, w = x.shape
    dev = _lib.require_cuda(x, lo, hi)
    _lib.no_autograd("channel_clamp", x)
    if lo.numel() != c or hi.numel() != c:  # broadcasting against the permuted [..., C] tensor
        raise RuntimeError(f"channel_clamp: bounds must have C = {c} entries, got {lo.numel()} / {hi.numel()}")
    x = x.contiguous()
    lo = lo.detach().to(torch.float32).contiguous()
    hi = hi.detach().to(torch.float32).contiguous()
    code = _lib.float_code(x)
    if out is None:
        out = torch.empty_like(x)
    elif out.shape != x.shape or out.dtype != x.dtype or not out.is_contiguous():
        raise ValueError("channel_clamp: `out` must be a contiguous tensor shaped and typed like x")
    if x.numel() > 0:
        with _lib.on_device(dev):
            st = _lib.load().udape_channel_clamp(x.data_ptr(), code, n * c, c, h * w, lo.data_ptr(), hi.data_ptr(),
                                                 out.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(st, "channel_clamp")
    return out

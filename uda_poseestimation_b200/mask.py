"""Teacher-confidence masks — the inline fragments of the trainers restated as functions:

* ``train_human.py:376-383`` (same in ``train_animal.py:401-408``)::

      conf = y_t_tea_recon.amax(dim=(2,3))
      pred_position = y_t_tea_recon.view(b, k, -1).argmax(-1)
      pred_position = torch.stack([pred_position % w, pred_position // w], -1)
      conf_table = conf >= args.occlude_thresh

* ``train_human.py:427-430`` (``train_animal.py:452-455``)::

      activates = y_t_tea_recon.amax(dim=(2,3))
      mask_thresh = torch.kthvalue(activates.view(-1), int(mask_ratio * activates.numel()))[0].item()
      tea_mask = tea_mask * activates > mask_thresh

The reference reads the teacher heatmaps three times (amax, argmax, amax again), runs a
``kthvalue`` and synchronises the host with ``.item()``.  Here the statistics come out of the
single decode launch and the k-th value is selected on the device (exact radix select), so
there is no host sync; ``mask_thresh`` stays a device scalar.
"""
from __future__ import annotations

import torch

from . import _lib
from .keypoint_detection import decode

__all__ = ["confidence_mask", "consistency_mask", "teacher_targets", "teacher_targets_rewarped"]


def confidence_mask(hm: torch.Tensor, occlude_thresh: float):
    """``(conf float32[B,K], pred_position int64[B,K,2] (x,y), conf_table bool[B,K])``."""
    r = decode(hm.detach(), want_maxvals=True, want_position=True, occlude_thresh=float(occlude_thresh))
    return r["maxvals"].squeeze(-1), r["position"], r["conf_table"]


def consistency_mask(activates: torch.Tensor, mask_ratio: float, tea_mask: torch.Tensor | None = None):
    """``(tea_mask bool[B,K], mask_thresh 0-dim float32 tensor)`` from per-joint activations.

    ``k = int(mask_ratio * activates.numel())`` is the 1-based rank passed to
    ``torch.kthvalue``; like torch, ``k`` outside ``[1, numel]`` raises.
    """
    dev = _lib.require_cuda(activates, tea_mask)
    act = activates.detach()
    if act.dtype != torch.float32:
        act = act.float()
    act = act.contiguous()
    n = act.numel()
    kth = int(mask_ratio * n)
    if not (1 <= kth <= n):
        raise IndexError(f"kthvalue(): selected number k out of range for dimension 0 (k={kth}, n={n})")
    tm_in = None
    if tea_mask is not None:
        if tea_mask.numel() != n:
            raise ValueError("consistency_mask: tea_mask must have one entry per activation")
        tm_in = tea_mask.detach().to(torch.float32).contiguous()
    out = torch.empty(act.shape, dtype=torch.bool, device=dev)
    thresh = torch.empty((), dtype=torch.float32, device=dev)
    with _lib.on_device(dev):
        st = _lib.load().udape_mask_select(act.data_ptr(), n, kth, _lib.ptr(tm_in), thresh.data_ptr(),
                                           out.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(st, "consistency_mask")
    return out, thresh


def teacher_targets(hm: torch.Tensor, sigma, mask_ratio: float, occlude_thresh: float | None = None,
                    tea_mask: torch.Tensor | None = None, materialise: bool = True) -> dict:
    """Everything the trainers derive from the reconstructed teacher heatmaps
    (train_human.py:376-383 and :427-430) in ONE launch: a fused decode(+rectify) pass over ``hm`` whose
    last CTA runs the k-th-value select (``udape_decode_select``).

    Returns ``activates`` float32[B,K], ``preds`` float32[B,K,2] (the arg-max ``rectify`` pastes its
    window at), ``rectified`` (= ``rectify(hm, sigma)``), ``tea_mask`` bool[B,K], ``mask_thresh``
    (device scalar) and, when ``occlude_thresh`` is given, ``conf`` (alias of activates),
    ``position`` int64[B,K,2] and ``conf_table`` bool[B,K].  With ``materialise=False`` the
    rectified map is not written (``rectified`` is None): ``fused_losses(..., tea_preds=preds,
    sigma=sigma)`` evaluates it on the fly.
    """
    kth = int(mask_ratio * hm.shape[0] * hm.shape[1])   # train_human.py:429
    r = decode(hm.detach(), want_preds=True, want_maxvals_f32=True, want_position=occlude_thresh is not None,
               occlude_thresh=occlude_thresh, rectify_sigma=float(sigma) if materialise else None,
               select_kth=kth, select_tea_mask=tea_mask)
    act = r["maxvals_f32"]
    mask, thresh = r["tea_mask"], r["mask_thresh"]
    out = {"activates": act, "preds": r["preds"], "rectified": r.get("rectified"), "tea_mask": mask,
           "mask_thresh": thresh}
    if occlude_thresh is not None:
        out.update(conf=act, position=r["position"], conf_table=r["conf_table"])
    return out


def teacher_targets_rewarped(y_t_tea: torch.Tensor, theta: torch.Tensor, sigma, mask_ratio: float,
                             occlude_thresh: float | None = None, tea_mask: torch.Tensor | None = None) -> dict:
    """``teacher_targets(gather(y_t_tea, theta), ..., materialise=False)`` — the whole teacher chain of the step for ONE
    teacher view (train_human.py:359-372, :376-383, :427-430) — in one launch where ``rewarp.gather_decode`` has one
    (64 x 64 heatmaps), in two otherwise.  The re-warped teacher map is only ever decoded, so it is not written:
    ``y_t_tea_recon`` is ``None`` in the result (``gather`` it when it is wanted).  Same keys as
    :func:`teacher_targets` otherwise, ``rectified`` is ``None`` (``fused_losses(..., tea_preds=preds, sigma=sigma)``)."""
    from . import rewarp as _rewarp

    y = y_t_tea.detach()
    if not _rewarp.gather_decode_supported(y.contiguous()):
        out = teacher_targets(_rewarp.gather(y, theta), sigma, mask_ratio, occlude_thresh=occlude_thresh, tea_mask=tea_mask,
                              materialise=False)
        out["y_t_tea_recon"] = None
        return out
    kth = int(mask_ratio * y.shape[0] * y.shape[1])   # train_human.py:429
    r = _rewarp.gather_decode(y, theta, want_preds=True, want_position=occlude_thresh is not None,
                              occlude_thresh=occlude_thresh, select_kth=kth, select_tea_mask=tea_mask)
    act = r["maxvals_f32"]
    out = {"activates": act, "preds": r["preds"], "rectified": None, "tea_mask": r["tea_mask"],
           "mask_thresh": r["mask_thresh"], "y_t_tea_recon": None}
    if occlude_thresh is not None:
        out.update(conf=act, position=r["position"], conf_table=r["conf_table"])
    return out

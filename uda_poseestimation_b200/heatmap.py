"""Gaussian heatmap writers — drop-in for ``lib/datasets/util.py`` ``generate_target``
(:12-70) and ``draw_labelmap_ori`` (:326-363), and ``utils.py`` ``rectify`` (:77-109).

The reference runs these per sample / per joint in numpy inside DataLoader workers
(``rendered_hand_pose_mt.py:99-147``) or, for ``rectify``, in a B×K Python loop with four
host syncs per joint.  The batched functions here write a whole ``[B,K,H,W]`` tensor with
one CUDA launch; the single-sample functions keep the reference signatures on top of them.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .keypoint_detection import decode

__all__ = ["generate_target", "generate_target_batched", "generate_targets_multi", "draw_labelmap_ori",
           "draw_labelmap_batched", "draw_labelmaps_multi", "rectify"]


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("uda_poseestimation_b200: no CUDA device; this package has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def generate_target_batched(joints, joints_vis, heatmap_size, sigma, image_size, device=None, out=None):
    """Batched ``generate_target``: ``joints[...,K,2]`` (image pixels), ``joints_vis[...,K,1]``
    → ``(target[...,K,H,W] float32, target_weight[...,K,1] float32)`` on the GPU.

    ``heatmap_size`` and ``image_size`` are ``(W, H)`` as in the reference.  Placement is
    evaluated in float64 (``mu = int(joint / stride + 0.5)``, util.py:38-39), so pass float64
    keypoints when they come from numpy.  ``out=(target, weight)``: write into existing contiguous float32
    CUDA tensors of those shapes (e.g. the static inputs of a captured step graph).
    """
    if isinstance(joints, np.ndarray):
        joints = torch.from_numpy(np.ascontiguousarray(joints, dtype=np.float64))
    if isinstance(joints_vis, np.ndarray):
        joints_vis = torch.from_numpy(np.ascontiguousarray(joints_vis, dtype=np.float32))
    if device is None:
        device = joints.device if joints.is_cuda else _default_device()
    device = torch.device(device)
    lead = tuple(joints.shape[:-1])
    if joints.shape[-1] != 2:
        raise ValueError(f"generate_target: joints must end in 2 coordinates, got {tuple(joints.shape)}")
    planes = int(np.prod(lead)) if lead else 1
    if joints_vis.numel() != planes and joints_vis.shape[:len(lead)] != lead:
        raise ValueError("generate_target: joints_vis must have one entry per joint")
    j = joints.detach().to(device=device, dtype=torch.float64).reshape(planes, 2).contiguous()
    v = joints_vis.detach().to(device=device, dtype=torch.float32).reshape(planes, -1)[:, 0].contiguous()
    hm_w, hm_h = int(heatmap_size[0]), int(heatmap_size[1])
    if out is not None:
        target, weight = out
        for t, shape in ((target, lead + (hm_h, hm_w)), (weight, lead + (1,))):
            if (tuple(t.shape) != shape or t.dtype != torch.float32 or not t.is_contiguous() or t.device != device):
                raise ValueError(f"generate_target: out tensors must be contiguous float32 {shape} on {device}")
    else:
        target = torch.empty(lead + (hm_h, hm_w), dtype=torch.float32, device=device)
        weight = torch.empty(lead + (1,), dtype=torch.float32, device=device)
    if planes > 0:
        with _lib.on_device(device):
            st = _lib.load().udape_gauss_target(j.data_ptr(), v.data_ptr(), planes, hm_w, hm_h, float(sigma),
                                                float(image_size[0]), float(image_size[1]),
                                                target.data_ptr(), weight.data_ptr(), _lib.stream_ptr(device))
        _lib.check(st, "generate_target")
    return target, weight


def generate_targets_multi(joint_sets, joints_vis, heatmap_sizes, sigma, image_size, device=None):
    """Every target set a loader builds per sample, for the whole batch, in ONE launch
    (``rendered_hand_pose_mt.py:99,103,115,134,147``: five ``generate_target`` calls per sample — student,
    un-augmented and teacher-view keypoints at ``(64, 64)``, student and teacher view again at ``(8, 8)``).

    ``joint_sets``: list of ``[B,K,2]`` keypoint arrays (image pixels), ``joints_vis [B,K,1]`` shared by all sets
    (the reference passes the same ``visible``), ``heatmap_sizes``: one ``(W, H)`` per set.  Returns a list of
    ``(target [B,K,H,W] float32, target_weight [B,K,1] float32)`` CUDA tensors, one pair per set."""
    if len(joint_sets) != len(heatmap_sizes) or not 1 <= len(joint_sets) <= 8:
        raise ValueError("generate_targets_multi: one heatmap size per keypoint set, 1..8 sets")
    if device is None:
        first = joint_sets[0]
        device = first.device if (torch.is_tensor(first) and first.is_cuda) else _default_device()
    device = torch.device(device)

    def as_dev(x, dt):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        return x.detach().to(device=device, dtype=dt)

    lead = tuple(joint_sets[0].shape[:-1])
    planes = int(np.prod(lead))
    v = as_dev(joints_vis, torch.float32).reshape(planes, -1)[:, 0].contiguous()
    # one [S, planes, 2] float64 upload for all sets
    js = torch.stack([as_dev(j, torch.float64).reshape(planes, 2) for j in joint_sets]).contiguous()
    jobs = (_lib.TargetJob * len(joint_sets))()
    outs = []
    for i, (w, h) in enumerate(heatmap_sizes):
        if tuple(joint_sets[i].shape[:-1]) != lead or joint_sets[i].shape[-1] != 2:
            raise ValueError("generate_targets_multi: every keypoint set must be [..., K, 2] with the same leading shape")
        target = torch.empty(lead + (int(h), int(w)), dtype=torch.float32, device=device)
        weight = torch.empty(lead + (1,), dtype=torch.float32, device=device)
        jobs[i].joints, jobs[i].vis = js[i].data_ptr(), v.data_ptr()
        jobs[i].target, jobs[i].weight = target.data_ptr(), weight.data_ptr()
        jobs[i].hm_w, jobs[i].hm_h = int(w), int(h)
        outs.append((target, weight))
    if planes > 0:
        with _lib.on_device(device):
            st = _lib.load().udape_gauss_target_multi(jobs, len(joint_sets), planes, float(sigma), float(image_size[0]),
                                                      float(image_size[1]), _lib.stream_ptr(device))
        _lib.check(st, "generate_targets_multi")
    return outs


def generate_target(joints, joints_vis, heatmap_size, sigma, image_size):
    """Reference signature (util.py:12): ``joints (K,2)``, ``joints_vis (K,1)`` numpy arrays →
    ``(target (K,H,W), target_weight (K,1))`` numpy float32.  Torch inputs return CUDA tensors."""
    was_numpy = isinstance(joints, np.ndarray)
    target, weight = generate_target_batched(joints, joints_vis, heatmap_size, sigma, image_size)
    if was_numpy:
        return target.cpu().numpy(), weight.cpu().numpy()
    return target, weight


_KINDS = {"Gaussian": 0, "Cauchy": 1}


def draw_labelmap_batched(pts, height, width, sigma, type="Gaussian", out=None):
    """Batched ``draw_labelmap_ori``: ``pts[...,>=2]`` (heatmap pixels) →
    ``(img[...,H,W] float32, vis[...] int32)``; rejected joints (window touching the border,
    util.py:337-340) get an all-zero plane and ``vis = 0``.  With ``out`` given, windows are
    overwritten in place into that existing image batch instead."""
    if type not in _KINDS:
        raise ValueError(f"draw_labelmap_ori: unknown type {type!r}")  # reference: g undefined → NameError
    if isinstance(pts, np.ndarray):
        pts = torch.from_numpy(pts)
    device = pts.device if pts.is_cuda else (out.device if out is not None else _default_device())
    lead = tuple(pts.shape[:-1])
    planes = int(np.prod(lead)) if lead else 1
    # util.py:332  pt = pt.to(torch.int32)  (truncation toward zero)
    p = pts.detach()[..., :2].to(device=device).to(torch.int32).reshape(planes, 2).contiguous()
    zero_fill = out is None
    if out is None:
        out = torch.empty(lead + (int(height), int(width)), dtype=torch.float32, device=device)
    else:
        if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != lead + (int(height), int(width)):
            raise ValueError("draw_labelmap_batched: `out` must be a contiguous float32 [...,H,W] tensor")
        _lib.require_cuda(out)
    vis = torch.empty(lead, dtype=torch.int32, device=device)
    if planes > 0:
        with _lib.on_device(device):
            st = _lib.load().udape_labelmap(p.data_ptr(), planes, int(height), int(width), float(sigma),
                                            _KINDS[type], 1 if zero_fill else 0, out.data_ptr(),
                                            vis.data_ptr(), _lib.stream_ptr(device))
        _lib.check(st, "draw_labelmap_ori")
    return out, vis


def draw_labelmaps_multi(pt_sets, height, width, sigma, type="Gaussian", gates=None, device=None):
    """The label maps of several views of a batch in ONE launch (``real_animal_all_mt.py:275-283,306-311``:
    ``draw_labelmap_ori`` per joint for the un-augmented, the student's and each teacher view's keypoints, inside
    ``if tpts[i, 1] > 0``).

    ``pt_sets``: list of ``[B,K,>=2]`` point arrays in heatmap pixels (what the reference passes as
    ``tpts[i] - 1``), ``gates``: optional list (one per set, or one shared array) of ``[B,K]`` booleans — the
    reference's ``if``; a gated-off joint keeps an all-zero plane and ``vis = 1`` so that
    ``target_weight *= vis`` leaves its weight alone.  Returns a list of ``(img [B,K,H,W] float32, vis [B,K] int32)``."""
    if type not in _KINDS:
        raise ValueError(f"draw_labelmap_ori: unknown type {type!r}")
    if not 1 <= len(pt_sets) <= 8:
        raise ValueError("draw_labelmaps_multi: 1..8 point sets")
    if device is None:
        first = pt_sets[0]
        device = first.device if (torch.is_tensor(first) and first.is_cuda) else _default_device()
    device = torch.device(device)
    lead = tuple(pt_sets[0].shape[:-1])
    planes = int(np.prod(lead))

    def as_t(x):
        return torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x

    # util.py:332  pt = pt.to(torch.int32)  (truncation toward zero), one upload for all sets
    ps = torch.stack([as_t(p).detach()[..., :2].to(torch.int32).reshape(planes, 2) for p in pt_sets]).to(device).contiguous()
    if gates is not None and not isinstance(gates, (list, tuple)):
        gates = [gates] * len(pt_sets)
    gs = None
    if gates is not None:
        gs = torch.stack([as_t(g).detach().reshape(planes).to(torch.uint8) for g in gates]).to(device).contiguous()
    jobs = (_lib.LabelmapJob * len(pt_sets))()
    outs = []
    for i in range(len(pt_sets)):
        if tuple(pt_sets[i].shape[:-1]) != lead:
            raise ValueError("draw_labelmaps_multi: every point set must have the same leading shape")
        img = torch.empty(lead + (int(height), int(width)), dtype=torch.float32, device=device)
        vis = torch.empty(lead, dtype=torch.int32, device=device)
        jobs[i].pts, jobs[i].gate = ps[i].data_ptr(), (gs[i].data_ptr() if gs is not None else None)
        jobs[i].img, jobs[i].vis_out = img.data_ptr(), vis.data_ptr()
        outs.append((img, vis))
    if planes > 0:
        with _lib.on_device(device):
            st = _lib.load().udape_labelmap_multi(jobs, len(pt_sets), planes, int(height), int(width), float(sigma),
                                                  _KINDS[type], _lib.stream_ptr(device))
        _lib.check(st, "draw_labelmaps_multi")
    return outs


def draw_labelmap_ori(img, pt, sigma, type="Gaussian"):
    """Reference signature (util.py:326): draws into a copy of ``img [H,W]`` and returns
    ``(img tensor, vis ∈ {0,1})``.  The returned image lives on the GPU."""
    if isinstance(img, np.ndarray):
        img = torch.from_numpy(img)
    if not isinstance(pt, torch.Tensor):
        pt = torch.as_tensor(pt)
    dev = img.device if img.is_cuda else _default_device()
    canvas = img.detach().to(device=dev, dtype=torch.float32).clone().contiguous()
    h, w = canvas.shape
    out, vis = draw_labelmap_batched(pt.reshape(1, -1), h, w, sigma, type, out=canvas.view(1, h, w))
    return out[0], int(vis.item())


def rectify(hm: torch.Tensor, sigma) -> torch.Tensor:
    """Teacher pseudo-label: per (b,k) plane, zeros + a unit-peak Gaussian at the arg-max
    (utils.py:77-109).  One fused decode+write launch; same dtype/shape as ``hm``."""
    if hm.dim() != 4:
        raise ValueError(f"rectify: expected [B,K,H,W], got {tuple(hm.shape)}")
    return decode(hm.detach(), rectify_sigma=float(sigma))["rectified"]

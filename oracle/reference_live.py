"""The hot path through the reference's OWN functions — TEST INFRASTRUCTURE ONLY (never imported by the product).

Same call surface as ``oracle/reference_port.py``, but every name that IS a function or class of the reference
dispatches to the reference's code object (imported by ``oracle/ref_loader.py`` from ``/root/reference`` in the
build container, or from the bytecode ``oracle/build_ref.py`` compiled into ``oracle/_ref/`` on the GPU box):

    calc_mean_std, adaptive_instance_normalization      adain/function.py:3-22
    adain (inside adain_mix)                            lib/models/Style_net.py:21-29
    get_max_preds, calc_dists, dist_acc, accuracy       lib/keypoint_detection.py:9-94
    get_max_preds_torch, rectify, OldWeightEMA          utils.py:54-109, :9-25
    JointsMSELoss, ConsLoss                             lib/models/loss.py:11-49, :119-132
    generate_target, draw_labelmap_ori                  lib/datasets/util.py:12-70, :326-363
    ModelEMA                                            lib/models/ema.py:6-44

What the reference only has INLINE in its trainers is not importable and stays restated (the names below fall
through to the port, which ``tests/test_oracle_vs_reference.py`` pins): the alpha mix expression
(``Style_net.py:168``), the per-sample ``tF.affine`` re-warp loops (``train_human.py:359-372, 417-423`` — they call
torchvision, like the reference), the confidence / consistency masks (``:376-383, 427-430``).  The optimizer tail
(``:436-438, 441``) is the reference's own objects: ``torch.optim.Adam`` + ``GradScaler`` + ``OldWeightEMA``.

``bench.py`` uses this module for ``cpu_baseline`` / ``--impl reference`` (``kind: "reference"``) when
``ref_loader.available()``; ``tests/test_gpu_vs_reference.py`` uses it as the checker of the CUDA path.
"""
from __future__ import annotations

import torch

from oracle import ref_loader as _L
from oracle.reference_port import *  # noqa: F401,F403  (trainer-inline fragments; overridden below where the reference has a callable)
from oracle.reference_port import (confidence_mask, consistency_mask, student_recon, teacher_recon)  # noqa: F401


def available() -> bool:
    return _L.available()


def source() -> str | None:
    """"source" (reference tree) | "bytecode" (oracle/_ref) | None."""
    return _L.kind()


def calc_mean_std(feat, eps=1e-5):
    return _L.load("function").calc_mean_std(feat, eps)


def adaptive_instance_normalization(content_feat, style_feat):
    return _L.load("function").adaptive_instance_normalization(content_feat, style_feat)


def adain_mix(content_feat, style_feat, alpha=1.0):
    t = _L.load("style_net").adain(content_feat, style_feat)   # Style_net.py:167
    return alpha * t + (1 - alpha) * content_feat               # :168, an expression inside Net.forward


def get_max_preds(batch_heatmaps):
    return _L.load("keypoint_detection").get_max_preds(batch_heatmaps)


def get_max_preds_torch(batch_heatmaps):
    return _L.load("utils").get_max_preds_torch(batch_heatmaps)


def calc_dists(preds, target, normalize):
    return _L.load("keypoint_detection").calc_dists(preds, target, normalize)


def dist_acc(dists, thr=0.5):
    return _L.load("keypoint_detection").dist_acc(dists, thr)


def accuracy(output, target, hm_type="gaussian", thr=0.5):
    return _L.load("keypoint_detection").accuracy(output, target, hm_type, thr)


def joints_mse_loss(output, target, target_weight=None, reduction="mean"):
    return _L.load("loss").JointsMSELoss(reduction=reduction)(output, target, target_weight)


def cons_loss(stu_out, tea_out, valid_mask=None, tea_mask=None):
    return _L.load("loss").ConsLoss()(stu_out, tea_out, valid_mask=valid_mask, tea_mask=tea_mask)


def rectify(hm, sigma):
    return _L.load("utils").rectify(hm, sigma)


def generate_target(joints, joints_vis, heatmap_size, sigma, image_size):
    return _L.load("dataset_util").generate_target(joints, joints_vis, heatmap_size, sigma, image_size)


def draw_labelmap_ori(img, pt, sigma, type="Gaussian"):
    return _L.load("dataset_util").draw_labelmap_ori(img, pt, sigma, type=type)


class _ParamNet(torch.nn.Module):
    """A parameter list with the reference networks' census (the convolutions themselves are out of scope)."""

    def __init__(self, tensors):
        super().__init__()
        self.p = torch.nn.ParameterList([torch.nn.Parameter(t.detach().clone()) for t in tensors])


class TrainerTail:
    """``scaler.step(stu_optimizer); tea_optimizer.step(); scaler.update()`` (train_human.py:436-441) with the
    reference's own objects: ``torch.optim.Adam`` (:139), ``GradScaler`` (:324), ``OldWeightEMA`` (:141, utils.py:9-25).
    ``scaler`` is the one whose ``scale(loss).backward()`` ran this step (it owns the scale the gradients carry)."""

    def __init__(self, student_tensors, teacher_tensors, lr: float, alpha: float, loss_scale: float, device):
        dev = torch.device(device)
        self.student, self.teacher = _ParamNet(student_tensors).to(dev), _ParamNet(teacher_tensors).to(dev)
        for p in self.teacher.parameters():
            p.requires_grad_(False)
        self.opt = torch.optim.Adam(self.student.parameters(), lr=lr)                    # :139
        self.ema = _L.load("utils").OldWeightEMA(self.teacher, self.student, alpha=alpha)  # :141
        self.scaler = torch.amp.GradScaler(dev.type, init_scale=loss_scale)               # :324

    def step(self, scaled_grads):
        """``scaled_grads``: what ``scaler.scale(loss_all).backward()`` leaves in ``p.grad`` of the student."""
        for p, g in zip(self.student.parameters(), scaled_grads):
            p.grad = g.clone()
        self.scaler.step(self.opt)   # :436 (unscale + non-finite check + Adam)
        self.ema.step()              # :438
        self.scaler.update()         # :441

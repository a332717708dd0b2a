"""uda_poseestimation_b200 — B200-native (sm_100a) hot path of the mean-teacher + AdaIN
domain-adaptive pose trainer, behind the reference's own Python operator names.

Every operator is a hand-written CUDA kernel in ``csrc/`` reached through the C-ABI of
``include/udape.h`` (``libudape_b200.so``, loaded with ctypes).  There is no Triton, no
torch.compile, no CPU fallback: CPU tensors raise, and a missing shared library raises at
first use.

Reference name → module here
    adain/function.py, lib/models/Style_net.py : calc_mean_std, adaptive_instance_normalization, adain
                                                 (+ adain_mix = Style_net.py:167-168)
    adain/net.py:137-143                       : calc_style_loss (calc_mean_std with backward, planes up to 256x256)
    lib/keypoint_detection.py                  : get_max_preds, calc_dists, dist_acc, accuracy
    utils.py                                   : get_max_preds_torch, rectify, OldWeightEMA
    lib/models/loss.py                         : JointsMSELoss, ConsLoss
    lib/models/ema.py                          : ModelEMA
    lib/datasets/util.py                       : generate_target, draw_labelmap_ori
    train_human.py:376-383,427-430 (inline)    : confidence_mask, consistency_mask, teacher_targets, teacher_targets_rewarped
    train_human.py:359-372,417-423 (inline)    : teacher_recon, student_recon (three tF.affine calls per
                                                 sample → one gather launch, with backward)
    train_human.py:385-412 (inline)            : occlude_keypoints;  affine_nearest = batched tF.affine
    lib/models/Style_net.py:121-177 (as used)  : StyleTransfer (forward-only Net: encode, fused AdaIN+mix, decode; + clamp)
    train_human.py:136-141,260,436-440         : Adam, SGD, GradScaler (torch.optim / torch.cuda.amp drop-ins:
                                                 unscale + update + teacher EMA in one multi-tensor launch)
    train_human.py:145-148 (nn.DataParallel)   : PeerGroup, ShardedStudentStep (one process per GPU: gradient
                                                 reduce-scatter, sharded update and parameter all-gather + EMA
                                                 as peer-memory kernels over NVLink)
"""
from ._lib import UdapeError, check_tickets, library_path, load as load_library
from .adain import (adain, adain_mix, adain_mix_multi, adaptive_instance_normalization, calc_mean_std, calc_style_loss,
                    channel_clamp)
from .ema import ModelEMA, MultiTensorPlan, OldWeightEMA
from .heatmap import (draw_labelmap_batched, draw_labelmap_ori, draw_labelmaps_multi, generate_target,
                      generate_target_batched, generate_targets_multi, rectify)
from .keypoint_detection import (accuracy, accuracy_from_counts, calc_dists, decode, dist_acc, get_max_preds,
                                 get_max_preds_torch, pck_counts)
from .loss import ConsLoss, JointsMSELoss, cons_loss, fused_losses, joints_mse_loss
from .mask import confidence_mask, consistency_mask, teacher_targets, teacher_targets_rewarped
from .dp import PeerGroup, ShardedStudentStep
from .optim import SGD, Adam, GradScaler
from .stylize import StyleTransfer
from .rewarp import affine_nearest, occlude_keypoints, student_recon, teacher_recon

__version__ = "0.3.2"

__all__ = [
    "UdapeError", "library_path", "load_library", "check_tickets",
    "calc_mean_std", "calc_style_loss", "adaptive_instance_normalization", "adain", "adain_mix", "adain_mix_multi", "channel_clamp",
    "get_max_preds", "get_max_preds_torch", "calc_dists", "dist_acc", "accuracy", "pck_counts",
    "accuracy_from_counts", "decode",
    "JointsMSELoss", "ConsLoss", "joints_mse_loss", "cons_loss", "fused_losses",
    "generate_target", "generate_target_batched", "draw_labelmap_ori", "draw_labelmap_batched", "rectify",
    "generate_targets_multi", "draw_labelmaps_multi",
    "confidence_mask", "consistency_mask", "teacher_targets", "teacher_targets_rewarped",
    "OldWeightEMA", "ModelEMA", "MultiTensorPlan",
    "teacher_recon", "student_recon", "occlude_keypoints", "affine_nearest",
    "Adam", "SGD", "GradScaler", "StyleTransfer",
    "PeerGroup", "ShardedStudentStep",
]

"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A restatement, in plain torch / numpy on the CPU, of the reference's algorithm for every
function on the hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this module, and only as
the checker or the timed CPU baseline — never from the product package
(``uda_poseestimation_b200`` does not import it and has no CPU path).

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against the reference's *own functions executed in the build container*
(``tests/golden/make_golden.py`` imports them by file path from /root/reference and stores
seeded input/output fixtures in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
replays them, and ``tests/test_oracle_vs_reference.py`` compares live when the reference
tree is present).  Arithmetic that lives in un-vendored third-party code — torch 2.11 ATen
(var/mean/sqrt/mse_loss/argmax/amax/kthvalue/exp) and numpy 2.3 (argmax/amax/exp/
linalg.norm) — is called through the same public functions the reference calls.

Each function keeps the reference's op order (including its Python loops where the
reference loops) so that timing this module on the host is a fair "reference CPU path"
baseline.  Citations are file:line in the reference tree.
"""
from __future__ import annotations

import numpy as np
import torch

# --------------------------------------------------------------------------------------------------
# a1-a3  AdaIN  (adain/function.py:3-22, lib/models/Style_net.py:4-29,163-168)
# --------------------------------------------------------------------------------------------------


def calc_mean_std(feat, eps=1e-5):
    """function.py:3-11 — unbiased variance + eps, sqrt, then a separate mean pass."""
    assert feat.dim() == 4
    n, c = feat.shape[0], feat.shape[1]
    flat = feat.view(n, c, -1)
    std = (flat.var(dim=2) + eps).sqrt().view(n, c, 1, 1)
    mean = flat.mean(dim=2).view(n, c, 1, 1)
    return mean, std


def adaptive_instance_normalization(content_feat, style_feat):
    """function.py:14-22 — style statistics first, then content, expand-and-normalise."""
    assert content_feat.shape[:2] == style_feat.shape[:2]
    shape = content_feat.size()
    s_mean, s_std = calc_mean_std(style_feat)
    c_mean, c_std = calc_mean_std(content_feat)
    normalised = (content_feat - c_mean.expand(shape)) / c_std.expand(shape)
    return normalised * s_std.expand(shape) + s_mean.expand(shape)


def calc_style_loss(input, target):
    """adain/net.py:137-143 (``Net.calc_style_loss``): MSE of the channel statistics; autograd flows to
    ``input`` through calc_mean_std."""
    assert input.size() == target.size()
    assert target.requires_grad is False
    input_mean, input_std = calc_mean_std(input)
    target_mean, target_std = calc_mean_std(target)
    mse = torch.nn.MSELoss()
    return mse(input_mean, target_mean) + mse(input_std, target_std)


def style_transfer(encoder, decoder, content, style, alpha, recover_min=None, recover_max=None):
    """What the trainers keep of ``Style_net.Net.forward`` (Style_net.py:163-170; the content / Gram losses
    of :171-177 are discarded by every caller, train_human.py:275,350,355) followed by the clamp of :276.
    ``encoder`` is sliced like Net.__init__ (:122-126)."""
    enc = torch.nn.Sequential(*list(encoder.children())[:31])
    with torch.no_grad():
        style_feat = enc(style)
        content_feat = enc(content)
        t = adaptive_instance_normalization(content_feat, style_feat)
        t = alpha * t + (1 - alpha) * content_feat
        g_t = decoder(t)
        if recover_min is not None:
            g_t = channel_clamp(g_t, recover_min, recover_max)
    return g_t


def adain_mix(content_feat, style_feat, alpha=1.0):
    """Style_net.py:164,167-168 — t = adain(c, s); t = alpha * t + (1 - alpha) * c."""
    assert 0 <= alpha <= 1
    t = adaptive_instance_normalization(content_feat, style_feat)
    return alpha * t + (1 - alpha) * content_feat


def channel_clamp(x, recover_min, recover_max):
    """train_human.py:276,351,356 (train_animal.py:301,376,381) — the expression, verbatim."""
    return torch.maximum(torch.minimum(x.permute(0, 2, 3, 1), recover_max), recover_min).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------------------------------
# a7-a8  decode + PCK  (lib/keypoint_detection.py:9-94, utils.py:54-75)
# --------------------------------------------------------------------------------------------------


def get_max_preds(batch_heatmaps):
    """keypoint_detection.py:9-37 (numpy)."""
    assert isinstance(batch_heatmaps, np.ndarray)
    assert batch_heatmaps.ndim == 4
    b, k, _, w = batch_heatmaps.shape
    flat = batch_heatmaps.reshape((b, k, -1))
    idx = np.argmax(flat, 2).reshape((b, k, 1))
    maxvals = np.amax(flat, 2).reshape((b, k, 1))
    preds = np.tile(idx, (1, 1, 2)).astype(np.float32)
    preds[:, :, 0] = preds[:, :, 0] % w
    preds[:, :, 1] = np.floor(preds[:, :, 1] / w)
    keep = np.tile(np.greater(maxvals, 0.0), (1, 1, 2)).astype(np.float32)
    preds *= keep
    return preds, maxvals


def get_max_preds_torch(batch_heatmaps):
    """utils.py:54-75 (torch)."""
    b, k = batch_heatmaps.size(0), batch_heatmaps.size(1)
    w = batch_heatmaps.size(3)
    flat = batch_heatmaps.reshape((b, k, -1))
    idx = torch.argmax(flat, 2).reshape((b, k, 1))
    maxvals = torch.amax(flat, 2).reshape((b, k, 1))
    preds = idx.repeat(1, 1, 2).float()
    preds[:, :, 0] = preds[:, :, 0] % w
    preds[:, :, 1] = torch.floor(preds[:, :, 1] / w)
    keep = (maxvals > 0.0).repeat(1, 1, 2).float()
    preds *= keep
    return preds, maxvals


def calc_dists(preds, target, normalize):
    """keypoint_detection.py:40-52 — B×K Python loop, float64 distances, -1 for invalid."""
    preds = preds.astype(np.float32)
    target = target.astype(np.float32)
    dists = np.zeros((preds.shape[1], preds.shape[0]))
    for n in range(preds.shape[0]):
        for c in range(preds.shape[1]):
            if target[n, c, 0] > 1 and target[n, c, 1] > 1:
                a = preds[n, c, :] / normalize[n]
                t = target[n, c, :] / normalize[n]
                dists[c, n] = np.linalg.norm(a - t)
            else:
                dists[c, n] = -1
    return dists


def dist_acc(dists, thr=0.5):
    """keypoint_detection.py:55-62."""
    usable = np.not_equal(dists, -1)
    n = usable.sum()
    if n > 0:
        return np.less(dists[usable], thr).sum() * 1.0 / n
    return -1


def pck_counts(output, target, thr=0.5):
    """Integer (hits[K], valid[K]) behind ``accuracy`` — what the device kernel must match."""
    pred, _ = get_max_preds(output)
    tgt, _ = get_max_preds(target)
    h, w = output.shape[2], output.shape[3]
    norm = np.ones((pred.shape[0], 2)) * np.array([h, w]) / 10
    dists = calc_dists(pred, tgt, norm)
    usable = np.not_equal(dists, -1)
    hits = (np.less(dists, thr) & usable).sum(axis=1).astype(np.int32)
    return hits, usable.sum(axis=1).astype(np.int32), pred


def accuracy(output, target, hm_type="gaussian", thr=0.5):
    """keypoint_detection.py:65-94."""
    joints = list(range(output.shape[1]))
    norm = 1.0
    if hm_type == "gaussian":
        pred, _ = get_max_preds(output)
        target, _ = get_max_preds(target)
        h, w = output.shape[2], output.shape[3]
        norm = np.ones((pred.shape[0], 2)) * np.array([h, w]) / 10
    dists = calc_dists(pred, target, norm)
    acc = np.zeros(len(joints))
    avg_acc = 0
    cnt = 0
    for i in range(len(joints)):
        acc[i] = dist_acc(dists[joints[i]], thr)
        if acc[i] >= 0:
            avg_acc = avg_acc + acc[i]
            cnt += 1
    avg_acc = avg_acc / cnt if cnt != 0 else 0
    return acc, avg_acc, cnt, pred


# --------------------------------------------------------------------------------------------------
# a9-a10  losses  (lib/models/loss.py:11-49, :119-132)
# --------------------------------------------------------------------------------------------------


def joints_mse_loss(output, target, target_weight=None, reduction="mean"):
    """loss.py:39-49 — MSE(none) * 0.5, * weight.view(B,K,1), mean (all or per plane)."""
    b, k = output.shape[0], output.shape[1]
    pred = output.reshape((b, k, -1))
    gt = target.reshape((b, k, -1))
    loss = torch.nn.functional.mse_loss(pred, gt, reduction="none") * 0.5
    if target_weight is not None:
        loss = loss * target_weight.view((b, k, 1))
    if reduction == "mean":
        return loss.mean()
    if reduction == "none":
        return loss.mean(dim=-1)
    raise ValueError(reduction)


def cons_loss(stu_out, tea_out, valid_mask=None, tea_mask=None):
    """loss.py:124-132."""
    diff = stu_out - tea_out
    if tea_mask is not None:
        diff = diff * tea_mask[:, :, None, None]
    loss_map = torch.mean(diff ** 2, dim=1)
    if valid_mask is not None:
        loss_map = loss_map[valid_mask]
    return loss_map.mean()


# --------------------------------------------------------------------------------------------------
# a11  confidence masks  (train_human.py:376-383, 427-430 — inline fragments)
# --------------------------------------------------------------------------------------------------


def confidence_mask(hm, occlude_thresh):
    """train_human.py:377-383 → (conf[B,K], pred_position int64[B,K,2], conf_table bool[B,K])."""
    b, k, h, w = hm.size()
    conf = hm.amax(dim=(2, 3))
    flat_idx = hm.view(b, k, -1).argmax(-1)
    position = torch.stack([flat_idx % w, flat_idx // w], -1)
    return conf, position, conf >= occlude_thresh


def consistency_mask(hm, mask_ratio, tea_mask=None):
    """train_human.py:427,429-430 → (tea_mask bool[B,K], mask_thresh float, activates)."""
    activates = hm.amax(dim=(2, 3))
    if tea_mask is None:
        tea_mask = torch.ones_like(activates)  # train_human.py:360,372 sets it to all ones
    thresh = torch.kthvalue(activates.view(-1), int(mask_ratio * activates.numel()))[0].item()
    return tea_mask * activates > thresh, thresh, activates


# --------------------------------------------------------------------------------------------------
# a6  rectify  (utils.py:77-109)
# --------------------------------------------------------------------------------------------------


def rectify(hm, sigma):
    """utils.py:77-109 — B×K loop; note the reference checks x against h and y against w."""
    _, _, h, w = hm.size()
    out = torch.zeros_like(hm)
    coord, _ = get_max_preds_torch(hm)
    tmp_size = 3 * sigma
    for bi in range(out.size(0)):
        for ci in range(out.size(1)):
            mu_x = coord[bi, ci, 0]
            mu_y = coord[bi, ci, 1]
            ul = [int(mu_x - tmp_size), int(mu_y - tmp_size)]
            br = [int(mu_x + tmp_size + 1), int(mu_y + tmp_size + 1)]
            if mu_x >= h or mu_y >= w or mu_x < 0 or mu_y < 0:
                continue
            size = 2 * tmp_size + 1
            x = torch.arange(0, size, 1).float()
            y = x.unsqueeze(1)
            x0 = y0 = size // 2
            g = torch.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))
            g_x = max(0, -ul[0]), min(br[0], h) - ul[0]
            g_y = max(0, -ul[1]), min(br[1], w) - ul[1]
            img_x = max(0, ul[0]), min(br[0], h)
            img_y = max(0, ul[1]), min(br[1], w)
            out[bi][ci][img_y[0]:img_y[1], img_x[0]:img_x[1]] = g[g_y[0]:g_y[1], g_x[0]:g_x[1]]
    return out


# --------------------------------------------------------------------------------------------------
# a4-a5  target heatmaps  (lib/datasets/util.py:12-70, :326-363)
# --------------------------------------------------------------------------------------------------


def generate_target(joints, joints_vis, heatmap_size, sigma, image_size):
    """util.py:12-70 — joints (K,2), joints_vis (K,1); sizes are (W,H)."""
    k = joints.shape[0]
    weight = np.ones((k, 1), dtype=np.float32)
    weight[:, 0] = joints_vis[:, 0]
    target = np.zeros((k, heatmap_size[1], heatmap_size[0]), dtype=np.float32)
    tmp_size = sigma * 3
    image_size = np.array(image_size)
    heatmap_size = np.array(heatmap_size)
    for j in range(k):
        stride = image_size / heatmap_size
        mu_x = int(joints[j][0] / stride[0] + 0.5)
        mu_y = int(joints[j][1] / stride[1] + 0.5)
        ul = [int(mu_x - tmp_size), int(mu_y - tmp_size)]
        br = [int(mu_x + tmp_size + 1), int(mu_y + tmp_size + 1)]
        if mu_x >= heatmap_size[0] or mu_y >= heatmap_size[1] or mu_x < 0 or mu_y < 0:
            weight[j] = 0
            continue
        size = 2 * tmp_size + 1
        x = np.arange(0, size, 1, np.float32)
        y = x[:, np.newaxis]
        x0 = y0 = size // 2
        g = np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))
        g_x = max(0, -ul[0]), min(br[0], heatmap_size[0]) - ul[0]
        g_y = max(0, -ul[1]), min(br[1], heatmap_size[1]) - ul[1]
        img_x = max(0, ul[0]), min(br[0], heatmap_size[0])
        img_y = max(0, ul[1]), min(br[1], heatmap_size[1])
        if weight[j] > 0.5:
            target[j][img_y[0]:img_y[1], img_x[0]:img_x[1]] = g[g_y[0]:g_y[1], g_x[0]:g_x[1]]
    return target, weight


def draw_labelmap_ori(img, pt, sigma, type="Gaussian"):
    """util.py:326-363 — img [H,W] torch/numpy, pt tensor; returns (torch img, vis)."""
    img = img.numpy().copy() if torch.is_tensor(img) else np.array(img, copy=True)
    pt = pt.to(torch.int32)
    ul = [int(pt[0] - 3 * sigma), int(pt[1] - 3 * sigma)]
    br = [int(pt[0] + 3 * sigma + 1), int(pt[1] + 3 * sigma + 1)]
    if br[0] >= img.shape[1] or br[1] >= img.shape[0] or ul[0] < 0 or ul[1] < 0:
        return torch.from_numpy(img), 0
    size = 6 * sigma + 1
    x = np.arange(0, size, 1, float)
    y = x[:, np.newaxis]
    x0 = y0 = size // 2
    if type == "Gaussian":
        g = np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))
    elif type == "Cauchy":
        g = sigma / (((x - x0) ** 2 + (y - y0) ** 2 + sigma ** 2) ** 1.5)
    g_x = max(0, -ul[0]), min(br[0], img.shape[1]) - ul[0]
    g_y = max(0, -ul[1]), min(br[1], img.shape[0]) - ul[1]
    img_x = max(0, ul[0]), min(br[0], img.shape[1])
    img_y = max(0, ul[1]), min(br[1], img.shape[0])
    img[img_y[0]:img_y[1], img_x[0]:img_x[1]] = g[g_y[0]:g_y[1], g_x[0]:g_x[1]]
    return torch.from_numpy(img), 1


# --------------------------------------------------------------------------------------------------
# a12-a13  EMA  (utils.py:9-25, lib/models/ema.py:18-44)
# --------------------------------------------------------------------------------------------------


def ema_init(target_params, source_params):
    """utils.py:18-19."""
    for p, s in zip(target_params, source_params):
        p.data[:] = s.data[:]


def ema_step(target_params, source_params, alpha=0.999):
    """utils.py:21-25 — per-tensor mul_ then add_ of a scaled temporary."""
    one_minus_alpha = 1.0 - alpha
    for p, s in zip(target_params, source_params):
        p.data.mul_(alpha)
        p.data.add_(s.data * one_minus_alpha)


def model_ema_update(ema_tensors, model_tensors, decay, ema_buffers=(), model_buffers=()):
    """ema.py:22-36 on already-matched tensor lists."""
    with torch.no_grad():
        for e, m in zip(ema_tensors, model_tensors):
            e.copy_(e * decay + (1.0 - decay) * m)
        for e, m in zip(ema_buffers, model_buffers):
            e.copy_(m)


# --------------------------------------------------------------------------------------------------
# f2  student update + teacher EMA  (train_human.py:136-141, :436-440)
# --------------------------------------------------------------------------------------------------
# ``scaler.step(stu_optimizer); tea_optimizer.step(); scaler.update()``.  The optimizers are
# torch.optim.Adam(lr) / torch.optim.SGD(lr, momentum=0.9, weight_decay=1e-4, nesterov=True)
# (un-vendored torch; this image has 2.11).  Restated below from torch/optim/adam.py
# ``_single_tensor_adam``, torch/optim/sgd.py ``_single_tensor_sgd`` and torch/amp/grad_scaler.py
# (``_unscale_grads_`` / ``_maybe_opt_step``); pinned against those classes run on the CPU together
# with the reference's OldWeightEMA by tests/golden/make_golden.py (fixture ``optim.npz``).


def unscale_and_check(grads, scale):
    """GradScaler.unscale_: inv_scale = scale.double().reciprocal().float(); grad *= inv_scale;
    found_inf = any non-finite element (checked on the scaled gradient, like the ATen kernel)."""
    inv_scale = torch.as_tensor(scale, dtype=torch.float32).double().reciprocal().float()
    found_inf = False
    out = []
    for g in grads:
        if g is None:
            out.append(None)
            continue
        found_inf = found_inf or not bool(torch.isfinite(g).all())
        out.append(g * inv_scale)
    return out, found_inf


def adam_step(params, grads, exp_avgs, exp_avg_sqs, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """torch/optim/adam.py::_single_tensor_adam (amsgrad=False, maximize=False); ``step`` is the
    1-based number of this update.  In place on params / exp_avgs / exp_avg_sqs."""
    beta1, beta2 = betas
    for param, grad, exp_avg, exp_avg_sq in zip(params, grads, exp_avgs, exp_avg_sqs):
        if grad is None:
            continue
        if weight_decay != 0:
            grad = grad.add(param, alpha=weight_decay)
        exp_avg.lerp_(grad, 1 - beta1)
        exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        bias_correction1 = 1 - beta1 ** step
        bias_correction2 = 1 - beta2 ** step
        step_size = lr / bias_correction1
        bias_correction2_sqrt = bias_correction2 ** 0.5
        denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
        param.addcdiv_(exp_avg, denom, value=-step_size)


def sgd_step(params, grads, bufs, step, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
    """torch/optim/sgd.py::_single_tensor_sgd; ``bufs[i]`` is the momentum buffer (ignored on the
    first update, where torch clones the gradient into it).  In place."""
    for i, (param, grad) in enumerate(zip(params, grads)):
        if grad is None:
            continue
        if weight_decay != 0:
            grad = grad.add(param, alpha=weight_decay)
        if momentum != 0:
            buf = bufs[i]
            if step == 1:
                buf.copy_(grad)
            else:
                buf.mul_(momentum).add_(grad, alpha=1 - dampening)
            grad = grad.add(buf, alpha=momentum) if nesterov else buf
        param.add_(grad, alpha=-lr)


def student_teacher_step(algo, student, grads, state1, state2, teacher, step, scale, ema_alpha, **hyper):
    """train_human.py:436-438 on tensor lists: unscale + non-finite check, the optimizer update unless
    a gradient was non-finite (GradScaler._maybe_opt_step), then OldWeightEMA.step with the (possibly
    unchanged) student.  Returns (found_inf, step after the call)."""
    with torch.no_grad():
        if scale is not None:
            grads, found_inf = unscale_and_check(grads, scale)
        else:
            found_inf = False
        if not found_inf:
            step += 1
            if algo == "adam":
                adam_step(student, grads, state1, state2, step, **hyper)
            else:
                sgd_step(student, grads, state1, step, **hyper)
        ema_step(teacher, student, ema_alpha)
    return found_inf, step



def loader_targets_hand(kp_stu, kp_ori, kp_tea, visible, heatmap_size, sigma, image_size):
    """lib/datasets/rendered_hand_pose_mt.py:99,103,115,134,147 for a batch (k = 1 teacher view): the five
    generate_target calls of one __getitem__, per sample.  Returns [(target [B,K,H,W], weight [B,K,1])] x 5 in
    the reference's call order: stu, ori, small stu (8, 8), tea, small tea (8, 8)."""
    calls = [(kp_stu, heatmap_size), (kp_ori, heatmap_size), (kp_stu, (8, 8)), (kp_tea, heatmap_size), (kp_tea, (8, 8))]
    out = []
    for kp, size in calls:
        tg, wt = zip(*[generate_target(kp[i], visible[i], size, sigma, image_size) for i in range(kp.shape[0])])
        out.append((np.stack(tg), np.stack(wt)))
    return out


def loader_labelmaps_animal(pts, gate, weight0, res, sigma, type="Gaussian"):
    """lib/datasets/real_animal_all_mt.py:268-283 / :300-311 for one view of a batch: zero target, per joint
    `if gate: target[i], vis = draw_labelmap_ori(target[i], tpts[i] - 1, sigma, type); weight[i, 0] *= vis`.
    pts [B,K,>=2] (the 1-based `tpts`), gate [B,K] bool, weight0 [B,K] -> (target [B,K,res,res], weight [B,K,1])."""
    pts = torch.as_tensor(pts)
    b, k = pts.shape[:2]
    target = torch.zeros(b, k, res, res)
    weight = torch.as_tensor(weight0).clone().float().view(b, k, 1)
    for bi in range(b):
        for i in range(k):
            if bool(gate[bi][i]):
                target[bi, i], vis = draw_labelmap_ori(target[bi, i], pts[bi, i] - 1, sigma, type=type)
                weight[bi, i, 0] *= vis
    return target, weight

# --------------------------------------------------------------------------------------------------
# f1  affine re-warp loops  (train_human.py:359-372, :385-412, :418-423; same in train_animal.py)
# --------------------------------------------------------------------------------------------------
# The reference calls torchvision's ``tF.affine`` (an un-vendored, un-pinned dependency; this image
# has torchvision 0.26).  The loops below keep the reference's calls — per sample, three tF.affine
# calls, `.item()` on every parameter — so they ARE the reference semantics as long as torchvision
# is importable (it is in the build container and on the GPU box).  ``affine_nearest_restated``
# restates what one such call computes, in numpy, in the float32 op order of torchvision's
# `_gen_affine_grid` (CPU bmm = FMA chain over k) and ATen's `grid_sample(nearest)`; it is pinned
# bit-exactly against tF.affine by tests/test_oracle_golden.py and is what csrc/rewarp.cu follows.


def student_recon(y_t_stu, aug_param_stu, ratio, autocast=True):
    """train_human.py:417-423.  `autocast` reproduces the `torch.cuda.amp.autocast()` block on the
    CPU (`torch.autocast('cpu', float16)` has the same cast policy for bmm / grid_sampler)."""
    from torchvision.transforms import functional as tF

    angle, [trans_x, trans_y], [shear_x, shear_y], scale = aug_param_stu
    ctx = torch.autocast("cpu", dtype=y_t_stu.dtype) if (autocast and y_t_stu.dtype != torch.float32) else _null_ctx()
    with ctx:
        rows = []
        for ind in range(y_t_stu.size(0)):
            _angle, _trans_x, _trans_y, _shear_x, _shear_y, _scale = (angle[ind].item(), trans_x[ind].item(),
                                                                      trans_y[ind].item(), shear_x[ind].item(),
                                                                      shear_y[ind].item(), scale[ind].item())
            temp = tF.affine(y_t_stu[ind], 0., translate=[_trans_x / ratio, _trans_y / ratio], shear=[0., 0.], scale=1.)
            temp = tF.affine(temp, _angle, translate=[0., 0.], shear=[0., 0.], scale=_scale)
            # the reference assigns into zeros_like(y_t_stu): a cast back to the student dtype
            rows.append(tF.affine(temp, 0., translate=[0., 0.], shear=[_shear_x, _shear_y], scale=1.).to(y_t_stu.dtype))
        return torch.stack(rows, 0)


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def teacher_recon(y_t_teas, meta_aug_params, ratio):
    """train_human.py:359-372 — k views, CPU staging tensor `recons`, mean over views."""
    from torchvision.transforms import functional as tF

    k = len(y_t_teas)
    out = torch.zeros_like(y_t_teas[0])
    for ind in range(y_t_teas[0].size(0)):
        recons = torch.zeros(k, *y_t_teas[0].size()[1:])
        for _k in range(k):
            angle, [trans_x, trans_y], [shear_x, shear_y], scale = meta_aug_params[_k]
            _angle, _trans_x, _trans_y, _shear_x, _shear_y, _scale = (angle[ind].item(), trans_x[ind].item(),
                                                                      trans_y[ind].item(), shear_x[ind].item(),
                                                                      shear_y[ind].item(), scale[ind].item())
            temp = tF.affine(y_t_teas[_k][ind], 0., translate=[_trans_x / ratio, _trans_y / ratio], shear=[0., 0.], scale=1.)
            temp = tF.affine(temp, _angle, translate=[0., 0.], shear=[0., 0.], scale=_scale)
            temp = tF.affine(temp, 0., translate=[0, 0], shear=[_shear_x, _shear_y], scale=1.)
            recons[_k] = temp
        out[ind] = torch.mean(recons, dim=0)
    return out


def occlude_keypoints(x_t_stu, conf_table, pred_position, aug_param_stu, ratio, occlude_rate, occlude_size,
                      image_size, rng=np.random):
    """train_human.py:385-412 (with `np.int` spelled `int`: the alias is gone from numpy >= 1.24).
    pred_position is the host int array [B,K,2]; conf_table a bool tensor [B,K].  Returns a new
    tensor; a sample whose source patch overlaps the destination raises like the reference."""
    from torchvision.transforms import functional as tF

    x_t_stu = x_t_stu.clone()
    angle, [trans_x, trans_y], [shear_x, shear_y], scale = aug_param_stu
    b, k = conf_table.shape
    for _b in range(b):
        if conf_table[_b].sum() > 0 and rng.rand() <= occlude_rate:
            _angle, _trans_x, _trans_y, _shear_x, _shear_y, _scale = (angle[_b].item(), trans_x[_b].item(),
                                                                      trans_y[_b].item(), shear_x[_b].item(),
                                                                      shear_y[_b].item(), scale[_b].item())
            temp = tF.affine(x_t_stu[_b], 0., translate=[_trans_x / ratio, _trans_y / ratio], shear=[0., 0.], scale=1.)
            temp = tF.affine(temp, _angle, translate=[0., 0.], shear=[0., 0.], scale=_scale)
            temp = tF.affine(temp, 0., translate=[0., 0.], shear=[_shear_x, _shear_y], scale=1.)
            candidates = torch.arange(0, k)[conf_table[_b]]
            _c = rng.choice(candidates)
            position = (pred_position[_b, _c] * ratio).astype(int)
            left = max(position[1] - occlude_size, 0)
            right = min(position[1] + occlude_size, image_size)
            upper = max(position[0] - occlude_size, 0)
            bottom = min(position[0] + occlude_size, image_size)
            left_src = rng.randint(image_size - (right - left) + 1)
            right_src = left_src + right - left
            upper_src = rng.randint(image_size - (bottom - upper) + 1)
            bottom_src = upper_src + bottom - upper
            temp[:, left:right, upper:bottom] = temp[:, left_src:right_src, upper_src:bottom_src]
            x_t_stu[_b] = tF.affine(temp, -_angle, translate=[-_trans_x / ratio, -_trans_y / ratio],
                                    shear=[-_shear_x, -_shear_y], scale=1. / _scale)
    return x_t_stu


def inverse_affine_matrix(center, angle, translate, scale, shear):
    """torchvision 0.26 transforms/functional.py `_get_inverse_affine_matrix` (inverted=True),
    reached by the reference through tF.affine (train_human.py:366-368)."""
    import math

    rot, sx, sy = math.radians(angle), math.radians(shear[0]), math.radians(shear[1])
    cx, cy = center
    tx, ty = translate
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    matrix = [d, -b, 0.0, -c, a, 0.0]
    matrix = [x / scale for x in matrix]
    matrix[2] += matrix[0] * (-cx - tx) + matrix[1] * (-cy - ty)
    matrix[5] += matrix[3] * (-cx - tx) + matrix[4] * (-cy - ty)
    matrix[2] += cx
    matrix[5] += cy
    return matrix


def affine_source_index(matrix, height, width, theta_dtype=torch.float32, grid_dtype=None):
    """Source pixel of every output pixel of ONE tF.affine(nearest) call: int64 [H,W] flat index,
    -1 where grid_sample pads with zero.  torchvision `_gen_affine_grid` + ATen grid_sampler:

        theta = tensor(matrix, theta_dtype)            theta_dtype = dtype of the image tensor
        r = theta / (0.5*size) per row                 (in theta_dtype)
        x = linspace(-W/2+.5, W/2-.5, W), y likewise   (stored in theta_dtype)
        g = base_grid.bmm(r)                           under autocast bmm casts both operands to the
                                                       autocast dtype (= grid_dtype) and returns it;
                                                       CPU bmm = float32 FMA chain over k = 3:
                                                       fma(y, r1, x*r0) + r2, then one rounding
        ix = ((g + 1) * W - 1) / 2 in float32 (grid_sampler is autocast to float32);
        nearest = rint(ix); valid iff 0 <= nearest <= W-1
    """
    f32 = np.float32
    grid_dtype = grid_dtype or theta_dtype

    def rnd(a, dt):
        return a if dt == torch.float32 else torch.from_numpy(np.ascontiguousarray(a)).to(dt).float().numpy()

    theta = torch.tensor(matrix, dtype=theta_dtype).reshape(2, 3)
    r = (theta / torch.tensor([0.5 * width, 0.5 * height], dtype=theta_dtype).view(2, 1)).to(grid_dtype).float().numpy()
    x = rnd(rnd((np.arange(width, dtype=np.float64) + (0.5 - 0.5 * width)).astype(f32), theta_dtype), grid_dtype)
    y = rnd(rnd((np.arange(height, dtype=np.float64) + (0.5 - 0.5 * height)).astype(f32), theta_dtype), grid_dtype)
    xx, yy = np.meshgrid(x, y)
    src = []
    for row, size in ((0, width), (1, height)):
        t = (xx * r[row, 0]).astype(f32)                                   # rounded product
        t = (yy.astype(np.float64) * np.float64(r[row, 1]) + t.astype(np.float64)).astype(f32)  # fma (exact product)
        g = rnd((t + r[row, 2]).astype(f32), grid_dtype)
        i = ((g + f32(1)) * f32(size) - f32(1)) / f32(2)
        src.append(np.rint(i))
    ok = (src[0] >= 0) & (src[0] <= width - 1) & (src[1] >= 0) & (src[1] <= height - 1)
    flat = np.where(ok, src[1] * width + src[0], -1).astype(np.int64)
    return flat


def affine_nearest_restated(img, angle, translate, scale, shear, autocast_dtype=None):
    """One tF.affine(img [C,H,W], nearest) through `affine_source_index`; `autocast_dtype` = the
    dtype of an enclosing torch.autocast block (None: no autocast)."""
    c, h, w = img.shape
    m = inverse_affine_matrix([0.0, 0.0], angle, [1.0 * t for t in translate], scale, shear)
    src = torch.from_numpy(affine_source_index(m, h, w, img.dtype, autocast_dtype or img.dtype)).reshape(-1)
    flat = img.reshape(c, h * w)
    out = torch.where(src >= 0, flat[:, src.clamp_min(0)], torch.zeros((), dtype=img.dtype))
    return out.reshape(c, h, w)


def recon_source_index(angle, trans_x, trans_y, shear_x, shear_y, scale, ratio, height, width,
                       first_dtype=torch.float32, autocast_dtype=None):
    """Composed source index of the three-call chain (train_human.py:366-368 / :421-423) for one
    sample: out[p] = in[s1(s2(s3(p)))].  Under autocast the first call sees the half tensor (theta
    in half); grid_sample returns float32, so the later calls build float32 thetas — but bmm still
    casts them, and the base grid, to the autocast dtype: every grid is a half grid."""
    later = torch.float32 if autocast_dtype is not None else first_dtype
    calls = [(0.0, [trans_x / ratio, trans_y / ratio], 1.0, [0.0, 0.0], first_dtype),
             (angle, [0.0, 0.0], scale, [0.0, 0.0], later),
             (0.0, [0.0, 0.0], 1.0, [shear_x, shear_y], later)]
    maps = [affine_source_index(inverse_affine_matrix([0.0, 0.0], a, [1.0 * t for t in tr], sc, sh), height, width,
                                td, autocast_dtype or td).reshape(-1) for (a, tr, sc, sh, td) in calls]
    src = maps[2]                        # last applied call is evaluated first
    for m in (maps[1], maps[0]):
        src = np.where(src >= 0, m[np.clip(src, 0, None)], -1)
    return src


def recon_restated(y, aug_param, ratio, autocast_dtype=None):
    """student/teacher recon of one view through `recon_source_index` (forward), for pinning the
    composition + autocast rule that csrc/rewarp.cu implements."""
    angle, [trans_x, trans_y], [shear_x, shear_y], scale = aug_param
    b, c, h, w = y.shape
    out = torch.zeros_like(y)
    for ind in range(b):
        src = torch.from_numpy(recon_source_index(angle[ind].item(), trans_x[ind].item(), trans_y[ind].item(),
                                                  shear_x[ind].item(), shear_y[ind].item(), scale[ind].item(),
                                                  ratio, h, w, y.dtype, autocast_dtype))
        flat = y[ind].reshape(c, h * w)
        out[ind] = torch.where(src >= 0, flat[:, src.clamp_min(0)], torch.zeros((), dtype=y.dtype)).reshape(c, h, w)
    return out

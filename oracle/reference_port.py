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

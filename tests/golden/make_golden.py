"""Generates tests/golden/*.npz by running the REAL reference functions (imported by file
path from /root/reference, see oracle/ref_loader.py) on small seeded inputs.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

The fixtures store inputs AND the reference's outputs, so they are self-contained:
``tests/test_oracle_golden.py`` replays them against ``oracle/reference_port.py`` on the CPU
and ``tests/test_gpu_golden.py`` replays them against the CUDA operators on the GPU box.
The two inline trainer fragments that are not functions (train_human.py:376-383 and
:427-430) are executed here verbatim as expressions.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_loader  # noqa: E402
from uda_poseestimation_b200 import synthetic  # noqa: E402

OUT = Path(__file__).resolve().parent


def save(name, **arrays):
    clean = {}
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        clean[k] = np.asarray(v)
    np.savez_compressed(OUT / f"{name}.npz", **clean)
    size = (OUT / f"{name}.npz").stat().st_size
    print(f"{name}.npz  {size / 1024:.1f} KiB  keys={sorted(clean)}")


def golden_adain():
    fn = ref_loader.load("function")
    sn = ref_loader.load("style_net")
    out = {}
    cases = {"a": (2, 8, 32, 32, 32, 32), "b": (2, 3, 5, 7, 5, 7), "c": (1, 4, 16, 16, 8, 8), "d": (2, 2, 33, 31, 17, 3)}
    for tag, (n, c, hc, wc, hs, ws) in cases.items():
        g = torch.Generator().manual_seed(100 + ord(tag))
        content = torch.relu(torch.randn(n, c, hc, wc, generator=g) + 0.2)
        style = torch.relu(torch.randn(n, c, hs, ws, generator=g) * 2 + 0.5)
        mean, std = fn.calc_mean_std(content)
        out[f"{tag}_content"], out[f"{tag}_style"] = content, style
        out[f"{tag}_mean"], out[f"{tag}_std"] = mean, std
        t = fn.adaptive_instance_normalization(content, style)
        assert torch.equal(t, sn.adain(content, style))
        out[f"{tag}_adain"] = t
        for alpha in (0.0, 0.37, 1.0):
            # Style_net.py:167-168
            t2 = sn.adain(content, style)
            t2 = alpha * t2 + (1 - alpha) * content
            out[f"{tag}_mix_{alpha}"] = t2
    save("adain", **out)


def golden_decode():
    kd = ref_loader.load("keypoint_detection")
    ut = ref_loader.load("utils")
    out = {}
    hm = synthetic.heatmaps(2, 4, seed=11, peak=(0.2, 1.1))
    adv = synthetic.adversarial_heatmaps(k=3, h=16, w=16, seed=5)
    odd = synthetic.heatmaps(2, 3, seed=12, h=9, w=13)
    for tag, t in (("hm", hm), ("adv", adv), ("odd", odd)):
        for dt_name, dt in (("f32", torch.float32), ("f16", torch.float16)):
            x = t.to(dt)
            preds, maxvals = kd.get_max_preds(x.numpy())
            out[f"{tag}_{dt_name}_in"] = x.numpy()
            out[f"{tag}_{dt_name}_preds"] = preds
            out[f"{tag}_{dt_name}_maxvals"] = maxvals
            if dt == torch.float32:
                p2, m2 = ut.get_max_preds_torch(x)
                out[f"{tag}_{dt_name}_preds_torch"] = p2
                out[f"{tag}_{dt_name}_maxvals_torch"] = m2
            flat = x.float().view(x.shape[0], x.shape[1], -1)
            out[f"{tag}_{dt_name}_idx"] = torch.argmax(flat, 2).to(torch.int32)
    save("decode", **out)


def golden_accuracy():
    kd = ref_loader.load("keypoint_detection")
    du = ref_loader.load("dataset_util")
    out = {}
    b, k = 6, 5
    joints, vis = synthetic.keypoints(b, k, seed=21)
    target = np.stack([du.generate_target(joints[i], vis[i], (64, 64), 2, (256, 256))[0] for i in range(b)])
    # predictions: the label shifted by a few pixels + noise, so that hits and misses both occur
    g = torch.Generator().manual_seed(22)
    shift = torch.randint(-4, 5, (b, k, 2), generator=g)
    pred = torch.zeros(b, k, 64, 64)
    tt = torch.from_numpy(target)
    for i in range(b):
        for j in range(k):
            pred[i, j] = torch.roll(tt[i, j], shifts=(int(shift[i, j, 1]), int(shift[i, j, 0])), dims=(0, 1))
    pred = pred + 0.02 * torch.randn(b, k, 64, 64, generator=g)
    for tag, o in (("f32", pred.numpy()), ("f16", pred.half().numpy())):
        acc, avg_acc, cnt, p = kd.accuracy(o, target)
        out[f"{tag}_output"] = o
        out[f"{tag}_acc"] = acc
        out[f"{tag}_avg_acc"] = np.float64(avg_acc)
        out[f"{tag}_cnt"] = np.int64(cnt)
        out[f"{tag}_pred"] = p
        # integer counts behind the ratios
        tp, _ = kd.get_max_preds(target)
        norm = np.ones((b, 2)) * np.array([64, 64]) / 10
        d = kd.calc_dists(p, tp, norm)
        out[f"{tag}_dists"] = d
        out[f"{tag}_hits"] = ((d != -1) & (d < 0.5)).sum(1).astype(np.int32)
        out[f"{tag}_valid"] = (d != -1).sum(1).astype(np.int32)
    out["target"] = target
    # adversarial: non-square, other threshold
    o2 = synthetic.heatmaps(3, 4, seed=23, h=48, w=32).numpy()
    t2 = synthetic.heatmaps(3, 4, seed=24, h=48, w=32, noise=0.0).numpy()
    acc, avg_acc, cnt, p = kd.accuracy(o2, t2, thr=1.5)
    out.update(ns_output=o2, ns_target=t2, ns_acc=acc, ns_avg_acc=np.float64(avg_acc), ns_cnt=np.int64(cnt), ns_pred=p)
    save("accuracy", **out)


def golden_losses():
    ls = ref_loader.load("loss")
    out = {}
    b, k = 3, 4
    g = torch.Generator().manual_seed(31)
    output = synthetic.heatmaps(b, k, seed=32, h=32, w=32)
    target = synthetic.heatmaps(b, k, seed=33, h=32, w=32, noise=0.0)
    weight = (torch.rand(b, k, 1, generator=g) > 0.3).float()
    weight[0, 0, 0] = 0.4
    out.update(output=output, target=target, weight=weight)
    for red in ("mean", "none"):
        for wtag, w in (("w", weight), ("now", None)):
            o = output.clone().requires_grad_(True)
            loss = ls.JointsMSELoss(reduction=red)(o, target, w)
            upstream = torch.full_like(loss, 1.0) if red == "mean" else torch.rand(loss.shape, generator=g)
            loss.backward(upstream)
            out[f"mse_{red}_{wtag}_loss"] = loss
            out[f"mse_{red}_{wtag}_upstream"] = upstream
            out[f"mse_{red}_{wtag}_grad"] = o.grad
    tea = synthetic.heatmaps(b, k, seed=34, h=32, w=32, noise=0.0)
    tea_mask = torch.rand(b, k, generator=g) > 0.4
    valid_mask = torch.rand(b, 32, 32, generator=g) > 0.5
    out.update(tea=tea, tea_mask=tea_mask, valid_mask=valid_mask)
    for tag, kw in (("plain", {}), ("tm", dict(tea_mask=tea_mask)), ("vm", dict(valid_mask=valid_mask)),
                    ("tmvm", dict(tea_mask=tea_mask, valid_mask=valid_mask))):
        s = output.clone().requires_grad_(True)
        loss = ls.ConsLoss()(s, tea, **kw)
        loss.backward(torch.tensor(2.5))
        out[f"cons_{tag}_loss"] = loss
        out[f"cons_{tag}_grad"] = s.grad
    save("losses", **out)


def golden_masks():
    out = {}
    hm = synthetic.heatmaps(4, 6, seed=41, peak=(0.3, 1.2))
    hm[0, 0] = -hm[0, 0].abs() - 0.1  # an all-negative plane
    out["hm"] = hm
    occlude_thresh, mask_ratio = 0.9, 0.5
    y_t_tea_recon = hm
    # ---- train_human.py:376-383 (verbatim) ----
    b, k, h, w = y_t_tea_recon.size()
    conf = y_t_tea_recon.amax(dim=(2, 3))
    pred_position = y_t_tea_recon.view(b, k, -1).argmax(-1)
    pred_position = torch.stack([pred_position % w, pred_position // w], -1).cpu().numpy()
    conf_table = conf >= occlude_thresh
    # ---- train_human.py:360,372 then :427,429-430 (verbatim) ----
    tea_mask = torch.ones(y_t_tea_recon.shape[:2])
    activates = y_t_tea_recon.amax(dim=(2, 3))
    mask_thresh = torch.kthvalue(activates.view(-1), int(mask_ratio * activates.numel()))[0].item()
    tea_mask = tea_mask * activates > mask_thresh
    out.update(conf=conf, pred_position=pred_position, conf_table=conf_table, activates=activates,
               mask_thresh=np.float32(mask_thresh), tea_mask=tea_mask, occlude_thresh=np.float32(occlude_thresh),
               mask_ratio=np.float64(mask_ratio))
    # ties at the threshold: quantised activations
    q = (hm * 4).round() / 4
    act = q.amax(dim=(2, 3))
    for r in (0.25, 0.5, 0.9):
        th = torch.kthvalue(act.view(-1), int(r * act.numel()))[0].item()
        out[f"q_thresh_{r}"] = np.float32(th)
        out[f"q_mask_{r}"] = torch.ones_like(act) * act > th
    out["q_hm"] = q
    save("masks", **out)


def golden_rectify():
    ut = ref_loader.load("utils")
    out = {}
    hm = synthetic.heatmaps(2, 4, seed=51, peak=(0.3, 1.2))
    adv = synthetic.adversarial_heatmaps(k=2, h=32, w=32, seed=6)
    out.update(hm=hm, adv=adv)
    for sig_tag, sigma in (("2", 2), ("1.0", 1.0), ("1.5", 1.5)):
        out[f"hm_rect_{sig_tag}"] = ut.rectify(hm, sigma)
        out[f"adv_rect_{sig_tag}"] = ut.rectify(adv, sigma)
    save("rectify", **out)


def golden_targets():
    du = ref_loader.load("dataset_util")
    out = {}
    joints, vis = synthetic.keypoints(3, 8, seed=61)
    # corner cases: exactly on borders, negative fractions (int() truncates toward zero), weight 0.4
    joints[0, 0] = [-2.1, 10.0]
    joints[0, 1] = [255.9, 255.9]
    joints[0, 2] = [0.0, 0.0]
    joints[0, 3] = [254.0, 1.9]
    joints[0, 4] = [-1.9, -1.9]
    vis[0, 5, 0] = 0.4
    vis[0, 6, 0] = 0.6
    out.update(joints=joints, vis=vis)
    for tag, hs, sigma in (("64_s2", (64, 64), 2), ("64_s1", (64, 64), 1.0), ("8_s2", (8, 8), 2), ("48x32_s1", (48, 32), 1)):
        tg, wt = zip(*[du.generate_target(joints[i], vis[i], hs, sigma, (256, 256)) for i in range(joints.shape[0])])
        out[f"target_{tag}"] = np.stack(tg)
        out[f"weight_{tag}"] = np.stack(wt)
    # draw_labelmap_ori (animal variant)
    pts = torch.tensor([[10.7, 20.2], [2.0, 2.0], [3.0, 3.0], [60.0, 60.0], [61.0, 30.0], [30.0, 59.9], [31.5, 0.0],
                        [-4.0, 5.0], [33.0, 47.0]])
    out["lm_pts"] = pts
    for tag, sigma, kind in (("g1", 1.0, "Gaussian"), ("g2", 2, "Gaussian"), ("c1", 1.0, "Cauchy")):
        imgs, viss = [], []
        for p in pts:
            im, v = du.draw_labelmap_ori(torch.zeros(64, 64), p, sigma, type=kind)
            imgs.append(im.numpy())
            viss.append(v)
        out[f"lm_img_{tag}"] = np.stack(imgs)
        out[f"lm_vis_{tag}"] = np.array(viss, dtype=np.int32)
    # drawing into a non-zero canvas
    canvas = torch.full((64, 64), 0.25)
    im, v = du.draw_labelmap_ori(canvas, torch.tensor([20.0, 30.0]), 1.0)
    out["lm_canvas"] = im.numpy()
    save("targets", **out)


def golden_loader_targets():
    """The loader-side call patterns, verbatim: rendered_hand_pose_mt.py:99,103,115,134,147 (five generate_target
    calls per sample: three at heatmap_size, two at (8, 8)) and real_animal_all_mt.py:268-283,306-311 (three
    draw_labelmap_ori calls per joint inside `if tpts[i, 1] > 0`, weights multiplied by the returned vis)."""
    du = ref_loader.load("dataset_util")
    out = {}
    b, k = 4, 21
    rng = np.random.RandomState(71)
    visible = (rng.uniform(size=(b, k, 1)) >= 0.1).astype(np.float32)
    sets = {n: synthetic.keypoints(b, k, seed=72 + i)[0] for i, n in enumerate(("stu", "ori", "tea"))}
    sets["stu"][0, 0] = [-2.1, 300.0]
    sets["tea"][1, 3] = [255.9, 0.4]
    hs, sigma, image = (64, 64), 2, (256, 256)
    for n, kp in sets.items():
        out[f"hand_kp_{n}"] = kp
    out["hand_visible"] = visible
    calls = [("stu", hs), ("ori", hs), ("stu", (8, 8)), ("tea", hs), ("tea", (8, 8))]    # :99, :103, :115, :134, :147
    for ci, (n, size) in enumerate(calls):
        tg, wt = zip(*[du.generate_target(sets[n][i], visible[i], size, sigma, image) for i in range(b)])
        out[f"hand_target_{ci}"], out[f"hand_weight_{ci}"] = np.stack(tg), np.stack(wt)
    # animal variant: pts [K,3] = (x, y, visible) per sample, already in heatmap pixels + 1 (the reference's
    # `transform(...)` output is 1-based; it passes tpts[i] - 1)
    nparts, res, sig = 18, 64, 1.0
    for kind in ("Gaussian", "Cauchy"):
        views = {v: rng.uniform(-3.0, 68.0, size=(b, nparts, 3)).astype(np.float32) for v in ("ori", "stu", "tea")}
        for v in views.values():
            v[..., 2] = (rng.uniform(size=(b, nparts)) >= 0.15)
            v[rng.uniform(size=(b, nparts)) < 0.1, 1] = -1.0      # `if tpts[i, 1] > 0` skips these
        # the reference gates all three draws of :280-281 on the STUDENT's y and :310 on the teacher's y
        gate = {"ori": views["stu"][..., 1] > 0, "stu": views["stu"][..., 1] > 0, "tea": views["tea"][..., 1] > 0}
        for v, pts in views.items():
            target = torch.zeros(b, nparts, res, res)
            weight = torch.tensor(views["stu" if v != "tea" else "tea"][..., 2]).clone().view(b, nparts, 1)
            tp = torch.tensor(pts)
            for bi in range(b):
                for i in range(nparts):
                    if gate[v][bi, i]:
                        target[bi, i], vis = du.draw_labelmap_ori(target[bi, i], tp[bi, i] - 1, sig, type=kind)
                        weight[bi, i, 0] *= vis
            out[f"animal_{kind}_pts_{v}"], out[f"animal_{kind}_gate_{v}"] = pts, gate[v]
            out[f"animal_{kind}_w0_{v}"] = views["stu" if v != "tea" else "tea"][..., 2]
            out[f"animal_{kind}_target_{v}"], out[f"animal_{kind}_weight_{v}"] = target, weight
    save("loader_targets", **out)


def golden_ema():
    ut = ref_loader.load("utils")
    em = ref_loader.load("ema")
    out = {}

    def make(seed):
        torch.manual_seed(seed)
        return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 5, 1),
                                   torch.nn.Linear(7, 3))

    teacher, student = make(1), make(2)
    out["student0"] = np.concatenate([p.detach().numpy().ravel() for p in student.parameters()])
    out["teacher0"] = np.concatenate([p.detach().numpy().ravel() for p in teacher.parameters()])
    opt = ut.OldWeightEMA(teacher, student, alpha=0.999)
    out["teacher_init"] = np.concatenate([p.detach().numpy().ravel() for p in teacher.parameters()])
    g = torch.Generator().manual_seed(3)
    deltas = []
    for step in range(3):
        d = [torch.randn(p.shape, generator=g) * 0.05 for p in student.parameters()]
        with torch.no_grad():
            for p, dd in zip(student.parameters(), d):
                p.add_(dd)
        deltas.append(np.concatenate([x.numpy().ravel() for x in d]))
        opt.step()
        out[f"teacher_step{step}"] = np.concatenate([p.detach().numpy().ravel() for p in teacher.parameters()])
    out["deltas"] = np.stack(deltas)
    # ModelEMA (its __init__ calls .cuda(); make that a no-op in this GPU-less container)
    orig_cuda = torch.nn.Module.cuda
    torch.nn.Module.cuda = lambda self, device=None: self
    try:
        model = make(4)
        mema = em.ModelEMA(model, decay=0.99)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
            model[1].running_mean.add_(1.0)
            model[1].num_batches_tracked.add_(3)
        out["mema_model"] = np.concatenate([p.detach().numpy().ravel() for p in model.parameters()])
        out["mema_before"] = np.concatenate([p.detach().numpy().ravel() for p in mema.ema.parameters()])
        mema.update(model)
        out["mema_after"] = np.concatenate([p.detach().numpy().ravel() for p in mema.ema.parameters()])
        out["mema_buffers_after"] = np.concatenate([b.detach().double().numpy().ravel() for b in mema.ema.buffers()])
        mema.momentum_update(model, 0.9)
        out["mema_after_momentum"] = np.concatenate([p.detach().numpy().ravel() for p in mema.ema.parameters()])
    finally:
        torch.nn.Module.cuda = orig_cuda
    save("ema", **out)


def golden_optim():
    """train_human.py:136-141 + :436-440 on the CPU: torch.optim.Adam / SGD(nesterov) + the reference's
    OldWeightEMA driven through torch.amp.GradScaler (scaled gradients are assigned directly; one step
    carries an inf and must be skipped by scaler.step while the EMA still runs)."""
    ut = ref_loader.load("utils")
    out = {}

    def make(seed):
        torch.manual_seed(seed)
        return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 5, 1),
                                   torch.nn.Linear(7, 3))

    def flat(ts):
        return np.concatenate([t.detach().numpy().ravel() for t in ts])

    configs = {
        "adam": lambda ps: torch.optim.Adam(ps, lr=1e-3),                               # train_human.py:139
        "adamwd": lambda ps: torch.optim.Adam(ps, lr=3e-4, betas=(0.8, 0.99), eps=1e-6, weight_decay=1e-2),
        "sgd": lambda ps: torch.optim.SGD(ps, lr=0.1, momentum=0.9, weight_decay=0.0001, nesterov=True),  # :137
        "sgdplain": lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.8, dampening=0.1),
    }
    n_steps, bad_step = 5, 2
    for tag, mk in configs.items():
        teacher, student = make(1), make(2)
        opt = mk(student.parameters())
        tea = ut.OldWeightEMA(teacher, student, alpha=0.99)
        scaler = torch.amp.GradScaler("cpu", init_scale=1024.0, growth_interval=2)
        out[f"{tag}_student0"] = flat(student.parameters())
        g = torch.Generator().manual_seed(11)
        for step in range(n_steps):
            scale = float(scaler.scale(torch.ones(())))   # initialises / reads the current scale
            grads = [torch.randn(p.shape, generator=g) * 0.1 * scale for p in student.parameters()]
            if step == bad_step:
                grads[2].view(-1)[3] = float("inf")
            for p, gr in zip(student.parameters(), grads):
                p.grad = gr.clone()
            out[f"{tag}_scale{step}"] = np.float32(scale)
            out[f"{tag}_grads{step}"] = flat(grads)
            scaler.step(opt)       # :437
            tea.step()             # :438
            scaler.update()        # :440
            out[f"{tag}_student{step + 1}"] = flat(student.parameters())
            out[f"{tag}_teacher{step + 1}"] = flat(teacher.parameters())
        st = [opt.state[p] for p in student.parameters()]
        if tag.startswith("adam"):
            out[f"{tag}_exp_avg"] = flat([x["exp_avg"] for x in st])
            out[f"{tag}_exp_avg_sq"] = flat([x["exp_avg_sq"] for x in st])
            out[f"{tag}_steps"] = np.float32(st[0]["step"].item())
        else:
            out[f"{tag}_momentum_buffer"] = flat([x["momentum_buffer"] for x in st])
        out[f"{tag}_final_scale"] = np.float32(scaler.get_scale())
    save("optim", **out)


def golden_style_loss():
    """adain/net.py:137-143 called on the real class (unbound, with the module's own MSELoss) + torch
    autograd through the reference's calc_mean_std: loss and d loss / d input."""
    import types
    net = ref_loader.load("adain_net")
    fake_self = types.SimpleNamespace(mse_loss=torch.nn.MSELoss())
    out = {}
    cases = {"warp": (2, 3, 8, 8), "ragged": (1, 2, 5, 7), "stream": (1, 2, 40, 40), "mid": (2, 2, 32, 32)}
    for tag, shape in cases.items():
        g = torch.Generator().manual_seed(300 + len(tag))
        x = torch.relu(torch.randn(*shape, generator=g) + 0.3).requires_grad_(True)
        t = torch.relu(torch.randn(*shape, generator=g) * 1.5 + 0.1)
        loss = net.Net.calc_style_loss(fake_self, x, t)
        (gx,) = torch.autograd.grad(loss * 100.0, (x,))
        out[f"{tag}_input"], out[f"{tag}_target"], out[f"{tag}_loss"], out[f"{tag}_grad"] = x.detach(), t, loss.detach(), gx
        # the statistics' own backward with arbitrary upstream gradients
        fn = ref_loader.load("function")
        x2 = x.detach().clone().requires_grad_(True)
        m, s_ = fn.calc_mean_std(x2)
        dm, ds = torch.randn(m.shape, generator=g), torch.randn(s_.shape, generator=g)
        (g2,) = torch.autograd.grad([m, s_], (x2,), [dm, ds])
        out[f"{tag}_dmean"], out[f"{tag}_dstd"], out[f"{tag}_dfeat"] = dm, ds, g2
    save("style_loss", **out)


def golden_clamp():
    """train_human.py:276 verbatim, with both trainers' recover_min/max constants (:32-33, train_animal.py:34-35)."""
    out = {}
    bounds = {"human": ([-2.1179, -2.0357, -1.8044], [2.2489, 2.4285, 2.64]),
              "animal": ([-0.3999, -0.3909, -0.3871], [0.6001, 0.6091, 0.6129])}
    for tag, (lo, hi) in bounds.items():
        g = torch.Generator().manual_seed(700 + len(tag))
        x = torch.randn(3, 3, 17, 20, generator=g) * 2.0
        x[0, 1, 2, 3] = float("nan")
        x[1, 0, 0, 0] = float("inf")
        x[2, 2, 5, 5] = -float("inf")
        recover_min, recover_max = torch.tensor(lo), torch.tensor(hi)
        y = torch.maximum(torch.minimum(x.permute(0, 2, 3, 1), recover_max), recover_min).permute(0, 3, 1, 2)
        out[f"{tag}_x"], out[f"{tag}_lo"], out[f"{tag}_hi"], out[f"{tag}_y"] = x, recover_min, recover_max, y.contiguous()
    save("clamp", **out)


def golden_rewarp():
    """train_human.py:359-372 (teacher recon), :417-423 (student recon under autocast, with autograd)
    and :385-412 (occlusion) executed as written, through torchvision's tF.affine on the CPU.  The
    fragments are inline trainer code, so they are restated here line by line; `.cuda()` calls are
    dropped and `torch.cuda.amp.autocast()` becomes `torch.autocast('cpu', float16)` (same cast
    policy: bmm stays half, grid_sampler runs in float32)."""
    from torchvision.transforms import functional as tF

    out = {}
    ratio = 256 / 64

    def flat_aug(tag, ap):
        angle, (tx, ty), (sx, sy), sc = ap
        out[f"{tag}_aug"] = np.stack([angle.numpy(), tx.numpy().astype(np.float64), ty.numpy().astype(np.float64),
                                      sx.numpy(), sy.numpy(), sc.numpy()], 1)

    # ---- teacher recon, k = 1 and k = 2 views, float32 --------------------------------------------
    for tag, k, shape in (("tea1", 1, (3, 4, 64, 64)), ("tea2", 2, (2, 3, 64, 64)), ("teaodd", 1, (2, 5, 24, 40))):
        y_t_teas = [synthetic.heatmaps(shape[0], shape[1], seed=70 + i, h=shape[2], w=shape[3], peak=(0.3, 1.2)) for i in range(k)]
        meta_t_tea = [{"aug_param_tea": synthetic.aug_params(shape[0], seed=80 + i, shear_y=(tag == "teaodd"))} for i in range(k)]
        y_t_tea_recon = torch.zeros_like(y_t_teas[0])
        for ind in range(y_t_teas[0].size(0)):
            recons = torch.zeros(k, *y_t_teas[0].size()[1:])
            for _k in range(k):
                angle, [trans_x, trans_y], [shear_x, shear_y], scale = meta_t_tea[_k]["aug_param_tea"]
                _angle, _trans_x, _trans_y, _shear_x, _shear_y, _scale = angle[ind].item(), trans_x[ind].item(), trans_y[ind].item(), shear_x[ind].item(), shear_y[ind].item(), scale[ind].item()
                temp = tF.affine(y_t_teas[_k][ind], 0., translate=[_trans_x / ratio, _trans_y / ratio], shear=[0., 0.], scale=1.)
                temp = tF.affine(temp, _angle, translate=[0., 0.], shear=[0., 0.], scale=_scale)
                temp = tF.affine(temp, 0., translate=[0, 0], shear=[_shear_x, _shear_y], scale=1.)
                recons[_k] = temp
            y_t_tea_recon[ind] = torch.mean(recons, dim=0)
        for i in range(k):
            out[f"{tag}_in{i}"] = y_t_teas[i]
            flat_aug(f"{tag}_v{i}", meta_t_tea[i]["aug_param_tea"])
        out[f"{tag}_out"] = y_t_tea_recon

    # ---- student recon under autocast, with the gradient of sum(recon * G) --------------------------
    for tag, dt in (("stu16", torch.float16), ("stubf", torch.bfloat16), ("stu32", torch.float32)):
        y_t_stu = synthetic.heatmaps(3, 4, seed=90).to(dt).requires_grad_(True)
        aug = synthetic.aug_params(3, seed=91)
        angle, [trans_x, trans_y], [shear_x, shear_y], scale = aug
        G = torch.randn(3, 4, 64, 64, generator=torch.Generator().manual_seed(92))
        with torch.autocast("cpu", dtype=dt, enabled=dt != torch.float32):
            y_t_stu_recon = torch.zeros_like(y_t_stu)
            rows = []
            for ind in range(y_t_stu.size(0)):
                _angle, _trans_x, _trans_y, _shear_x, _shear_y, _scale = angle[ind].item(), trans_x[ind].item(), trans_y[ind].item(), shear_x[ind].item(), shear_y[ind].item(), scale[ind].item()
                temp = tF.affine(y_t_stu[ind], 0., translate=[_trans_x / ratio, _trans_y / ratio], shear=[0., 0.], scale=1.)
                temp = tF.affine(temp, _angle, translate=[0., 0.], shear=[0., 0.], scale=_scale)
                # `y_t_stu_recon[ind] = ...` on a zeros_like(y_t_stu) buffer: a cast to the student dtype
                rows.append(tF.affine(temp, 0., translate=[0., 0.], shear=[_shear_x, _shear_y], scale=1.).to(dt))
            y_t_stu_recon = torch.stack(rows, 0)
        (y_t_stu_recon.float() * G).sum().backward()
        out[f"{tag}_in"] = y_t_stu.detach().float()
        flat_aug(tag, aug)
        out[f"{tag}_G"] = G
        out[f"{tag}_out"] = y_t_stu_recon.detach().float()
        out[f"{tag}_grad"] = y_t_stu.grad.float()

    # ---- occlusion, :385-412 (np.int -> int), small images: image_size 64, heatmaps 16 -------------
    image_size, occlude_size, occlude_rate = 64, 6, 0.7
    b, k = 4, 5
    for seed in range(100, 200):
        x_t_stu = torch.randn(b, 3, image_size, image_size, generator=torch.Generator().manual_seed(seed))
        x_in = x_t_stu.clone()
        aug = synthetic.aug_params(b, seed=seed + 1, image=image_size)
        tea = synthetic.heatmaps(b, k, seed=seed + 2, h=16, w=16, peak=(0.5, 1.4))
        angle, [trans_x, trans_y], [shear_x, shear_y], scale = aug
        bb, kk, h, w = tea.size()
        conf = tea.amax(dim=(2, 3))
        pred_position = tea.view(bb, kk, -1).argmax(-1)
        pred_position = torch.stack([pred_position % w, pred_position // w], -1).cpu().numpy()
        conf_table = conf >= 0.9
        np.random.seed(seed)
        try:
            n_occluded = 0
            for _b in range(bb):
                if (conf_table[_b].sum() > 0 and np.random.rand() <= occlude_rate):
                    _angle, _trans_x, _trans_y, _shear_x, _shear_y, _scale = angle[_b].item(), trans_x[_b].item(), trans_y[_b].item(), shear_x[_b].item(), shear_y[_b].item(), scale[_b].item()
                    temp = tF.affine(x_t_stu[_b], 0., translate=[_trans_x / ratio, _trans_y / ratio], shear=[0., 0.], scale=1.)
                    temp = tF.affine(temp, _angle, translate=[0., 0.], shear=[0., 0.], scale=_scale)
                    temp = tF.affine(temp, 0., translate=[0., 0.], shear=[_shear_x, _shear_y], scale=1.)
                    candidates = torch.arange(0, kk)[conf_table[_b]]
                    _c = np.random.choice(candidates)
                    position = (pred_position[_b, _c] * ratio).astype(int)
                    left = max(position[1] - occlude_size, 0)
                    right = min(position[1] + occlude_size, image_size)
                    upper = max(position[0] - occlude_size, 0)
                    bottom = min(position[0] + occlude_size, image_size)
                    left_src = np.random.randint(image_size - (right - left) + 1)
                    right_src = left_src + right - left
                    upper_src = np.random.randint(image_size - (bottom - upper) + 1)
                    bottom_src = upper_src + bottom - upper
                    temp[:, left:right, upper:bottom] = temp[:, left_src:right_src, upper_src:bottom_src]
                    x_t_stu[_b] = tF.affine(temp, -_angle, translate=[-_trans_x / ratio, -_trans_y / ratio], shear=[-_shear_x, -_shear_y], scale=1. / _scale)
                    n_occluded += 1
        except RuntimeError:
            continue  # overlapping source / destination patch: torch refuses the copy, try another seed
        if 2 <= n_occluded < bb:
            break
    else:
        raise SystemExit("no occlusion seed found")
    out["occ_seed"] = np.int64(seed)
    out["occ_in"], out["occ_out"] = x_in, x_t_stu
    out["occ_conf_table"], out["occ_pred_position"] = conf_table, pred_position
    flat_aug("occ", aug)
    out["occ_args"] = np.array([ratio, occlude_rate, occlude_size, image_size], dtype=np.float64)
    save("rewarp", **out)


def main():
    if not ref_loader.available():
        raise SystemExit(f"reference tree not found at {ref_loader.REFERENCE_ROOT}")
    torch.manual_seed(0)
    np.random.seed(0)
    fns = (golden_adain, golden_decode, golden_accuracy, golden_losses, golden_masks, golden_rectify,
           golden_targets, golden_loader_targets, golden_ema, golden_optim, golden_style_loss, golden_clamp, golden_rewarp)
    only = set(sys.argv[1:])  # e.g. `make_golden.py clamp` regenerates one fixture
    for fn in fns:
        if not only or fn.__name__.removeprefix("golden_") in only:
            fn()


if __name__ == "__main__":
    main()

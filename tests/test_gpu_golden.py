"""GPU parity against the golden fixtures generated from the REAL reference
(tests/golden/make_golden.py).  Every call goes through the C-ABI (ctypes → libudape_b200.so).

Bars (BASELINE.json north star): bit-exact for indices / coordinates / masks / PCK counts /
integer placement; 1e-5 relative (scale-aware, see conftest.assert_close_scaled) for fp32
AdaIN outputs, heatmaps, losses and gradients; EMA fp32 is bit-exact by construction.
"""
import numpy as np
import pytest
import torch

import uda_poseestimation_b200 as U
from conftest import assert_close_scaled

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def C(a, dev):
    return torch.from_numpy(np.asarray(a)).to(dev)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_adain_golden(golden, dev, tag):
    g = golden("adain")
    c, s = C(g[f"{tag}_content"], dev), C(g[f"{tag}_style"], dev)
    mean, std = U.calc_mean_std(c)
    assert mean.shape == g[f"{tag}_mean"].shape and mean.dtype == torch.float32
    assert_close_scaled(mean, g[f"{tag}_mean"], RTOL, "mean")
    assert_close_scaled(std, g[f"{tag}_std"], RTOL, "std")
    assert_close_scaled(U.adaptive_instance_normalization(c, s), g[f"{tag}_adain"], RTOL, "adain")
    assert_close_scaled(U.adain(c, s), g[f"{tag}_adain"], RTOL, "adain alias")
    for alpha in (0.0, 0.37, 1.0):
        assert_close_scaled(U.adain_mix(c, s, alpha), g[f"{tag}_mix_{alpha}"], RTOL, f"mix {alpha}")
        a_dev = torch.tensor([alpha], dtype=torch.float32, device=dev)
        assert_close_scaled(U.adain_mix(c, s, a_dev), g[f"{tag}_mix_{alpha}"], RTOL, f"mix dev {alpha}")


@pytest.mark.parametrize("tag", ["hm", "adv", "odd"])
@pytest.mark.parametrize("dt", ["f32", "f16"])
def test_decode_golden(golden, dev, tag, dt):
    g = golden("decode")
    x = g[f"{tag}_{dt}_in"]
    # numpy in -> numpy out (reference contract)
    preds, maxvals = U.get_max_preds(x)
    assert isinstance(preds, np.ndarray) and preds.dtype == np.float32 and maxvals.dtype == x.dtype
    np.testing.assert_array_equal(preds, g[f"{tag}_{dt}_preds"])
    np.testing.assert_array_equal(maxvals, g[f"{tag}_{dt}_maxvals"])
    # tensor in -> tensor out, plus raw indices
    xt = C(x, dev)
    p2, m2 = U.get_max_preds_torch(xt)
    np.testing.assert_array_equal(p2.cpu().numpy(), g[f"{tag}_{dt}_preds"])
    np.testing.assert_array_equal(m2.cpu().numpy(), g[f"{tag}_{dt}_maxvals"])
    r = U.decode(xt, want_idx=True, want_position=True)
    np.testing.assert_array_equal(r["idx"].cpu().numpy(), g[f"{tag}_{dt}_idx"])
    w = x.shape[3]
    np.testing.assert_array_equal(r["position"][..., 0].cpu().numpy(), g[f"{tag}_{dt}_idx"] % w)
    np.testing.assert_array_equal(r["position"][..., 1].cpu().numpy(), g[f"{tag}_{dt}_idx"] // w)


@pytest.mark.parametrize("dt", ["f32", "f16"])
def test_accuracy_golden(golden, dev, dt):
    g = golden("accuracy")
    for as_numpy in (True, False):
        o = g[f"{dt}_output"] if as_numpy else C(g[f"{dt}_output"], dev)
        t = g["target"] if as_numpy else C(g["target"], dev)
        acc, avg_acc, cnt, pred = U.accuracy(o, t)
        np.testing.assert_array_equal(acc, g[f"{dt}_acc"])
        assert avg_acc == float(g[f"{dt}_avg_acc"]) and cnt == int(g[f"{dt}_cnt"])
        np.testing.assert_array_equal(pred, g[f"{dt}_pred"])
    hits, valid, _ = U.pck_counts(C(g[f"{dt}_output"], dev), C(g["target"], dev))
    np.testing.assert_array_equal(hits.cpu().numpy(), g[f"{dt}_hits"])
    np.testing.assert_array_equal(valid.cpu().numpy(), g[f"{dt}_valid"])


def test_accuracy_nonsquare_thr_golden(golden, dev):
    g = golden("accuracy")
    acc, avg_acc, cnt, pred = U.accuracy(g["ns_output"], g["ns_target"], thr=1.5)
    np.testing.assert_array_equal(acc, g["ns_acc"])
    assert avg_acc == float(g["ns_avg_acc"]) and cnt == int(g["ns_cnt"])
    np.testing.assert_array_equal(pred, g["ns_pred"])


@pytest.mark.parametrize("red", ["mean", "none"])
@pytest.mark.parametrize("wtag", ["w", "now"])
def test_joints_mse_golden(golden, dev, red, wtag):
    g = golden("losses")
    o = C(g["output"], dev).requires_grad_(True)
    w = C(g["weight"], dev) if wtag == "w" else None
    loss = U.JointsMSELoss(reduction=red)(o, C(g["target"], dev), w)
    assert loss.shape == g[f"mse_{red}_{wtag}_loss"].shape and loss.dtype == torch.float32
    assert_close_scaled(loss.detach(), g[f"mse_{red}_{wtag}_loss"], RTOL, "loss")
    loss.backward(C(g[f"mse_{red}_{wtag}_upstream"], dev))
    assert_close_scaled(o.grad, g[f"mse_{red}_{wtag}_grad"], RTOL, "grad")


@pytest.mark.parametrize("tag", ["plain", "tm", "vm", "tmvm"])
def test_cons_loss_golden(golden, dev, tag):
    g = golden("losses")
    s = C(g["output"], dev).requires_grad_(True)
    kw = {}
    if "tm" in tag:
        kw["tea_mask"] = C(g["tea_mask"], dev)
    if "vm" in tag:
        kw["valid_mask"] = C(g["valid_mask"], dev)
    loss = U.ConsLoss()(s, C(g["tea"], dev), **kw)
    assert_close_scaled(loss.detach(), g[f"cons_{tag}_loss"], RTOL, "loss")
    loss.backward(torch.tensor(2.5, device=dev))
    assert_close_scaled(s.grad, g[f"cons_{tag}_grad"], RTOL, "grad")


def test_masks_golden(golden, dev):
    g = golden("masks")
    hm = C(g["hm"], dev)
    conf, pos, table = U.confidence_mask(hm, float(g["occlude_thresh"]))
    np.testing.assert_array_equal(conf.cpu().numpy(), g["conf"])
    np.testing.assert_array_equal(pos.cpu().numpy(), g["pred_position"])
    np.testing.assert_array_equal(table.cpu().numpy(), g["conf_table"])
    assert table.dtype == torch.bool and pos.dtype == torch.int64
    mask, thresh = U.consistency_mask(C(g["activates"], dev), float(g["mask_ratio"]))
    np.testing.assert_array_equal(mask.cpu().numpy(), g["tea_mask"])
    assert np.float32(thresh.item()) == g["mask_thresh"]
    ones = torch.ones_like(C(g["activates"], dev))
    mask2, _ = U.consistency_mask(C(g["activates"], dev), float(g["mask_ratio"]), tea_mask=ones)
    np.testing.assert_array_equal(mask2.cpu().numpy(), g["tea_mask"])
    t = U.teacher_targets(hm, 2, float(g["mask_ratio"]), occlude_thresh=float(g["occlude_thresh"]))
    np.testing.assert_array_equal(t["tea_mask"].cpu().numpy(), g["tea_mask"])
    np.testing.assert_array_equal(t["conf_table"].cpu().numpy(), g["conf_table"])
    np.testing.assert_array_equal(t["activates"].cpu().numpy(), g["activates"])
    for r in (0.25, 0.5, 0.9):  # ties at the threshold are all excluded (strict >)
        q = C(g["q_hm"], dev)
        act = U.decode(q, want_maxvals_f32=True)["maxvals_f32"]
        mask, thresh = U.consistency_mask(act, r)
        np.testing.assert_array_equal(mask.cpu().numpy(), g[f"q_mask_{r}"])
        assert np.float32(thresh.item()) == g[f"q_thresh_{r}"]


@pytest.mark.parametrize("which", ["hm", "adv"])
@pytest.mark.parametrize("sig", [("2", 2), ("1.0", 1.0), ("1.5", 1.5)])
def test_rectify_golden(golden, dev, which, sig):
    g = golden("rectify")
    out = U.rectify(C(g[which], dev), sig[1])
    ref = g[f"{which}_rect_{sig[0]}"]
    assert out.dtype == torch.float32 and tuple(out.shape) == ref.shape
    # placement is exact: identical support and identical unit peaks
    np.testing.assert_array_equal(out.cpu().numpy() != 0, ref != 0)
    np.testing.assert_array_equal(out.cpu().numpy() == 1.0, ref == 1.0)
    assert_close_scaled(out, ref, RTOL, "rectify")


@pytest.mark.parametrize("case", [("64_s2", (64, 64), 2), ("64_s1", (64, 64), 1.0), ("8_s2", (8, 8), 2),
                                  ("48x32_s1", (48, 32), 1)])
def test_generate_target_golden(golden, dev, case):
    g = golden("targets")
    tag, hs, sigma = case
    # batched: [B,K,2] -> [B,K,H,W]
    t, w = U.generate_target_batched(g["joints"], g["vis"], hs, sigma, (256, 256))
    ref_t, ref_w = g[f"target_{tag}"], g[f"weight_{tag}"]
    np.testing.assert_array_equal(w.cpu().numpy(), ref_w)  # weights: exact
    np.testing.assert_array_equal(t.cpu().numpy() != 0, ref_t != 0)  # integer placement: exact
    np.testing.assert_array_equal(t.cpu().numpy() == 1.0, ref_t == 1.0)
    assert_close_scaled(t, ref_t, RTOL, "target")
    # reference signature: one sample, numpy in/out
    t1, w1 = U.generate_target(g["joints"][0], g["vis"][0], hs, sigma, (256, 256))
    assert isinstance(t1, np.ndarray) and t1.dtype == np.float32 and w1.shape == (g["joints"].shape[1], 1)
    np.testing.assert_array_equal(t1 != 0, ref_t[0] != 0)
    np.testing.assert_array_equal(w1, ref_w[0])


def test_loader_side_target_sets_in_one_launch(golden, dev):
    """SURVEY.md §8f-4, second half: the five generate_target calls per sample of rendered_hand_pose_mt.py:99-147
    (three 64x64, two 8x8) as ONE launch over the batch, and the gated draw_labelmap_ori triples of
    real_animal_all_mt.py:275-283,306-311 as one launch — against the fixture made by the reference's functions."""
    g = golden("loader_targets")
    sets = [g["hand_kp_stu"], g["hand_kp_ori"], g["hand_kp_stu"], g["hand_kp_tea"], g["hand_kp_tea"]]
    sizes = [(64, 64), (64, 64), (8, 8), (64, 64), (8, 8)]
    outs = U.generate_targets_multi(sets, g["hand_visible"], sizes, 2, (256, 256))
    assert len(outs) == 5
    for ci, (t, w) in enumerate(outs):
        ref_t, ref_w = g[f"hand_target_{ci}"], g[f"hand_weight_{ci}"]
        assert tuple(t.shape) == ref_t.shape and tuple(w.shape) == ref_w.shape
        np.testing.assert_array_equal(w.cpu().numpy(), ref_w)                    # weights: exact
        np.testing.assert_array_equal(t.cpu().numpy() != 0, ref_t != 0)          # integer placement: exact
        np.testing.assert_array_equal(t.cpu().numpy() == 1.0, ref_t == 1.0)
        assert_close_scaled(t, ref_t, RTOL, f"hand target set {ci}")
        # and equal to the single-set launch, bit for bit
        t1, w1 = U.generate_target_batched(sets[ci], g["hand_visible"], sizes[ci], 2, (256, 256))
        assert torch.equal(t, t1) and torch.equal(w, w1)
    for kind in ("Gaussian", "Cauchy"):
        views = ("ori", "stu", "tea")
        pts = [torch.from_numpy(g[f"animal_{kind}_pts_{v}"]) - 1 for v in views]      # the reference passes tpts[i] - 1
        gates = [g[f"animal_{kind}_gate_{v}"] for v in views]
        outs = U.draw_labelmaps_multi(pts, 64, 64, 1.0, kind, gates=gates)
        for v, (img, vis) in zip(views, outs):
            ref_t, ref_w = g[f"animal_{kind}_target_{v}"], g[f"animal_{kind}_weight_{v}"]
            np.testing.assert_array_equal(img.cpu().numpy() != 0, ref_t != 0)
            assert_close_scaled(img, ref_t, 1e-6, f"animal {kind} {v}")
            w = torch.from_numpy(g[f"animal_{kind}_w0_{v}"]).float().view(ref_w.shape) * vis.cpu().view(ref_w.shape)   # :282-283
            np.testing.assert_array_equal(w.numpy(), ref_w)


@pytest.mark.parametrize("case", [("g1", 1.0, "Gaussian"), ("g2", 2, "Gaussian"), ("c1", 1.0, "Cauchy")])
def test_draw_labelmap_golden(golden, dev, case):
    g = golden("targets")
    tag, sigma, kind = case
    img, vis = U.draw_labelmap_batched(g["lm_pts"], 64, 64, sigma, kind)
    np.testing.assert_array_equal(vis.cpu().numpy(), g[f"lm_vis_{tag}"])
    np.testing.assert_array_equal(img.cpu().numpy() != 0, g[f"lm_img_{tag}"] != 0)
    assert_close_scaled(img, g[f"lm_img_{tag}"], 1e-6, "labelmap")
    # reference signature, drawing into an existing canvas
    im, v = U.draw_labelmap_ori(torch.full((64, 64), 0.25), torch.tensor([20.0, 30.0]), 1.0)
    assert v == 1
    assert_close_scaled(im, g["lm_canvas"], 1e-6, "canvas")
    im0, v0 = U.draw_labelmap_ori(torch.zeros(64, 64), torch.tensor([2.0, 2.0]), 1.0)
    assert v0 == 0 and float(im0.abs().sum()) == 0.0


def _make(seed, dev):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 5, 1),
                               torch.nn.Linear(7, 3)).to(dev)


def _flat(params):
    return np.concatenate([p.detach().cpu().numpy().ravel() for p in params])


def _load_flat(module_params, flat):
    off = 0
    with torch.no_grad():
        for p in module_params:
            n = p.numel()
            p.copy_(torch.from_numpy(flat[off:off + n]).view(p.shape))
            off += n


def test_ema_golden(golden, dev):
    g = golden("ema")
    teacher, student = _make(1, dev), _make(2, dev)
    _load_flat(student.parameters(), g["student0"])
    _load_flat(teacher.parameters(), g["teacher0"])
    ptrs = [p.data_ptr() for p in teacher.parameters()]
    opt = U.OldWeightEMA(teacher, student, alpha=0.999)
    np.testing.assert_array_equal(_flat(teacher.parameters()), g["teacher_init"])
    for step in range(3):
        off = 0
        with torch.no_grad():
            for p in student.parameters():
                n = p.numel()
                p.add_(torch.from_numpy(g["deltas"][step][off:off + n]).view(p.shape).to(dev))
                off += n
        opt.step()
        # fp32 EMA is bit-identical to the eager reference (three roundings, no FMA)
        np.testing.assert_array_equal(_flat(teacher.parameters()), g[f"teacher_step{step}"])
    assert ptrs == [p.data_ptr() for p in teacher.parameters()], "EMA must update the live storages in place"
    # ModelEMA
    model = _make(4, dev)
    mema = U.ModelEMA(model, decay=0.99)
    _load_flat(model.parameters(), g["mema_model"])
    _load_flat(mema.ema.parameters(), g["mema_before"])
    with torch.no_grad():
        model[1].running_mean.add_(1.0)
        model[1].num_batches_tracked.add_(3)
    mema.update(model)
    np.testing.assert_array_equal(_flat(mema.ema.parameters()), g["mema_after"])
    np.testing.assert_array_equal(
        np.concatenate([b.detach().double().cpu().numpy().ravel() for b in mema.ema.buffers()]),
        g["mema_buffers_after"])
    mema.momentum_update(model, 0.9)
    np.testing.assert_array_equal(_flat(mema.ema.parameters()), g["mema_after_momentum"])


OPTIM_CASES = {
    "adam": (U.Adam, dict(lr=1e-3)),
    "adamwd": (U.Adam, dict(lr=3e-4, betas=(0.8, 0.99), eps=1e-6, weight_decay=1e-2)),
    "sgd": (U.SGD, dict(lr=0.1, momentum=0.9, weight_decay=0.0001, nesterov=True)),
    "sgdplain": (U.SGD, dict(lr=0.05, momentum=0.8, dampening=0.1)),
}


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("tag", sorted(OPTIM_CASES))
def test_student_step_golden(golden, dev, tag, fuse):
    """train_human.py:436-440 through the drop-in classes, in the reference's call order:
    scaler.step(stu_optimizer); tea_optimizer.step(); scaler.update().  Fixture = torch.optim + the
    reference's OldWeightEMA + torch.amp.GradScaler on the CPU, incl. one step with an inf gradient."""
    g = golden("optim")
    cls, kw = OPTIM_CASES[tag]
    teacher, student = _make(1, dev), _make(2, dev)
    _load_flat(student.parameters(), g[f"{tag}_student0"])
    opt = cls(student.parameters(), **kw)
    tea = U.OldWeightEMA(teacher, student, alpha=0.99)
    if fuse:
        opt.attach_teacher(tea)
    scaler = U.GradScaler(init_scale=1024.0, growth_interval=2)
    ptrs = [p.data_ptr() for p in list(student.parameters()) + list(teacher.parameters())]
    for p in student.parameters():
        p.grad = torch.zeros_like(p)
    for it in range(5):
        scale = float(scaler.scale(torch.ones((), device=dev)))
        assert scale == float(g[f"{tag}_scale{it}"])
        off = 0
        for p in student.parameters():
            n = p.numel()
            p.grad.copy_(torch.from_numpy(g[f"{tag}_grads{it}"][off:off + n]).view(p.shape))
            off += n
        scaler.step(opt)
        tea.step()
        scaler.update()
        assert_close_scaled(_flat(student.parameters()), g[f"{tag}_student{it + 1}"], RTOL, f"student after step {it}")
        assert_close_scaled(_flat(teacher.parameters()), g[f"{tag}_teacher{it + 1}"], RTOL, f"teacher after step {it}")
        if it == 2:  # the inf step: the student must be bit-identical to the previous state
            np.testing.assert_array_equal(_flat(student.parameters()), _prev)
        _prev = _flat(student.parameters())
    assert scaler.get_scale() == float(g[f"{tag}_final_scale"])
    assert opt.applied_steps() == 4
    assert ptrs == [p.data_ptr() for p in list(student.parameters()) + list(teacher.parameters())]
    st = [opt.state[p] for p in student.parameters()]
    if tag.startswith("adam"):
        assert_close_scaled(_flat([x["exp_avg"] for x in st]), g[f"{tag}_exp_avg"], RTOL, "exp_avg")
        assert_close_scaled(_flat([x["exp_avg_sq"] for x in st]), g[f"{tag}_exp_avg_sq"], RTOL, "exp_avg_sq")
        sd = opt.state_dict()
        assert float(sd["state"][0]["step"]) == float(g[f"{tag}_steps"])
    else:
        assert_close_scaled(_flat([x["momentum_buffer"] for x in st]), g[f"{tag}_momentum_buffer"], RTOL, "momentum_buffer")


@pytest.mark.parametrize("tag", ["warp", "ragged", "stream", "mid"])
def test_style_loss_golden(golden, dev, tag):
    """adain/net.py:137-143 with backward (decoder pre-training job): warp, three-sweep (ragged) and
    single-pass streaming statistics kernels + the elementwise backward."""
    g = golden("style_loss")
    x = C(g[f"{tag}_input"], dev).requires_grad_(True)
    loss = U.calc_style_loss(x, C(g[f"{tag}_target"], dev))
    (gx,) = torch.autograd.grad(loss * 100.0, (x,))
    assert_close_scaled(loss.detach(), g[f"{tag}_loss"], RTOL, "style loss")
    assert_close_scaled(gx, g[f"{tag}_grad"], RTOL, "d style loss / d input")
    x2 = C(g[f"{tag}_input"], dev).requires_grad_(True)
    m, s = U.calc_mean_std(x2)
    (g2,) = torch.autograd.grad([m, s], (x2,), [C(g[f"{tag}_dmean"], dev), C(g[f"{tag}_dstd"], dev)])
    assert_close_scaled(g2, g[f"{tag}_dfeat"], RTOL, "calc_mean_std backward")


@pytest.mark.parametrize("tag", ["human", "animal"])
def test_channel_clamp_golden(golden, dev, tag):
    g = golden("clamp")
    y = U.channel_clamp(C(g[f"{tag}_x"], dev), C(g[f"{tag}_lo"], dev), C(g[f"{tag}_hi"], dev))
    np.testing.assert_array_equal(y.cpu().numpy(), g[f"{tag}_y"])  # pure selection: bit-exact incl. NaN/Inf

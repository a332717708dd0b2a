"""Pins the CPU oracle (oracle/reference_port.py) against the golden fixtures that were
produced by the REAL reference functions (tests/golden/make_golden.py).  The oracle calls the
same torch/numpy primitives in the same order as the reference, so equality is exact."""
import numpy as np
import pytest
import torch

from conftest import assert_close_scaled
from oracle import reference_port as R


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_adain(golden, tag):
    g = golden("adain")
    c, s = T(g[f"{tag}_content"]), T(g[f"{tag}_style"])
    mean, std = R.calc_mean_std(c)
    assert torch.equal(mean, T(g[f"{tag}_mean"]))
    assert torch.equal(std, T(g[f"{tag}_std"]))
    assert torch.equal(R.adaptive_instance_normalization(c, s), T(g[f"{tag}_adain"]))
    for alpha in (0.0, 0.37, 1.0):
        assert torch.equal(R.adain_mix(c, s, alpha), T(g[f"{tag}_mix_{alpha}"]))


@pytest.mark.parametrize("tag", ["hm", "adv", "odd"])
@pytest.mark.parametrize("dt", ["f32", "f16"])
def test_get_max_preds(golden, tag, dt):
    g = golden("decode")
    x = g[f"{tag}_{dt}_in"]
    preds, maxvals = R.get_max_preds(x)
    np.testing.assert_array_equal(preds, g[f"{tag}_{dt}_preds"])
    np.testing.assert_array_equal(maxvals, g[f"{tag}_{dt}_maxvals"])  # NaN == NaN positionally
    assert preds.dtype == np.float32 and maxvals.dtype == x.dtype
    if dt == "f32":
        p2, m2 = R.get_max_preds_torch(T(x))
        np.testing.assert_array_equal(p2.numpy(), g[f"{tag}_{dt}_preds_torch"])
        np.testing.assert_array_equal(m2.numpy(), g[f"{tag}_{dt}_maxvals_torch"])


@pytest.mark.parametrize("dt", ["f32", "f16"])
def test_accuracy(golden, dt):
    g = golden("accuracy")
    acc, avg_acc, cnt, pred = R.accuracy(g[f"{dt}_output"], g["target"])
    np.testing.assert_array_equal(acc, g[f"{dt}_acc"])
    assert avg_acc == float(g[f"{dt}_avg_acc"]) and cnt == int(g[f"{dt}_cnt"])
    np.testing.assert_array_equal(pred, g[f"{dt}_pred"])
    hits, valid, _ = R.pck_counts(g[f"{dt}_output"], g["target"])
    np.testing.assert_array_equal(hits, g[f"{dt}_hits"])
    np.testing.assert_array_equal(valid, g[f"{dt}_valid"])


def test_accuracy_nonsquare_thr(golden):
    g = golden("accuracy")
    acc, avg_acc, cnt, pred = R.accuracy(g["ns_output"], g["ns_target"], thr=1.5)
    np.testing.assert_array_equal(acc, g["ns_acc"])
    assert avg_acc == float(g["ns_avg_acc"]) and cnt == int(g["ns_cnt"])
    np.testing.assert_array_equal(pred, g["ns_pred"])


@pytest.mark.parametrize("red", ["mean", "none"])
@pytest.mark.parametrize("wtag", ["w", "now"])
def test_joints_mse(golden, red, wtag):
    g = golden("losses")
    o = T(g["output"]).clone().requires_grad_(True)
    w = T(g["weight"]) if wtag == "w" else None
    loss = R.joints_mse_loss(o, T(g["target"]), w, red)
    assert torch.equal(loss.detach(), T(g[f"mse_{red}_{wtag}_loss"]))
    loss.backward(T(g[f"mse_{red}_{wtag}_upstream"]))
    assert torch.equal(o.grad, T(g[f"mse_{red}_{wtag}_grad"]))


@pytest.mark.parametrize("tag", ["plain", "tm", "vm", "tmvm"])
def test_cons_loss(golden, tag):
    g = golden("losses")
    s = T(g["output"]).clone().requires_grad_(True)
    kw = {}
    if "tm" in tag:
        kw["tea_mask"] = T(g["tea_mask"])
    if "vm" in tag:
        kw["valid_mask"] = T(g["valid_mask"])
    loss = R.cons_loss(s, T(g["tea"]), **kw)
    assert torch.equal(loss.detach(), T(g[f"cons_{tag}_loss"]))
    loss.backward(torch.tensor(2.5))
    assert torch.equal(s.grad, T(g[f"cons_{tag}_grad"]))


def test_masks(golden):
    g = golden("masks")
    hm = T(g["hm"])
    conf, pos, table = R.confidence_mask(hm, float(g["occlude_thresh"]))
    assert torch.equal(conf, T(g["conf"]))
    np.testing.assert_array_equal(pos.numpy(), g["pred_position"])
    assert torch.equal(table, T(g["conf_table"]))
    mask, thresh, act = R.consistency_mask(hm, float(g["mask_ratio"]))
    assert torch.equal(mask, T(g["tea_mask"])) and np.float32(thresh) == g["mask_thresh"]
    assert torch.equal(act, T(g["activates"]))
    for r in (0.25, 0.5, 0.9):
        mask, thresh, _ = R.consistency_mask(T(g["q_hm"]), r)
        assert torch.equal(mask, T(g[f"q_mask_{r}"])) and np.float32(thresh) == g[f"q_thresh_{r}"]


@pytest.mark.parametrize("which", ["hm", "adv"])
@pytest.mark.parametrize("sig", [("2", 2), ("1.0", 1.0), ("1.5", 1.5)])
def test_rectify(golden, which, sig):
    g = golden("rectify")
    out = R.rectify(T(g[which]), sig[1])
    assert torch.equal(out, T(g[f"{which}_rect_{sig[0]}"]))


@pytest.mark.parametrize("case", [("64_s2", (64, 64), 2), ("64_s1", (64, 64), 1.0), ("8_s2", (8, 8), 2),
                                  ("48x32_s1", (48, 32), 1)])
def test_generate_target(golden, case):
    g = golden("targets")
    tag, hs, sigma = case
    for i in range(g["joints"].shape[0]):
        t, w = R.generate_target(g["joints"][i], g["vis"][i], hs, sigma, (256, 256))
        np.testing.assert_array_equal(t, g[f"target_{tag}"][i])
        np.testing.assert_array_equal(w, g[f"weight_{tag}"][i])


@pytest.mark.parametrize("case", [("g1", 1.0, "Gaussian"), ("g2", 2, "Gaussian"), ("c1", 1.0, "Cauchy")])
def test_draw_labelmap(golden, case, capsys):
    g = golden("targets")
    tag, sigma, kind = case
    for i, p in enumerate(T(g["lm_pts"])):
        img, vis = R.draw_labelmap_ori(torch.zeros(64, 64), p, sigma, type=kind)
        np.testing.assert_array_equal(img.numpy(), g[f"lm_img_{tag}"][i])
        assert vis == int(g[f"lm_vis_{tag}"][i])
    img, vis = R.draw_labelmap_ori(torch.full((64, 64), 0.25), torch.tensor([20.0, 30.0]), 1.0)
    np.testing.assert_array_equal(img.numpy(), g["lm_canvas"])


def test_loader_side_target_sets(golden, capsys):
    """the loader call patterns (five generate_target calls per hand sample; gated draw_labelmap_ori triples per
    animal joint) restated by the oracle == the fixture produced by the reference's own functions"""
    g = golden("loader_targets")
    got = R.loader_targets_hand(g["hand_kp_stu"], g["hand_kp_ori"], g["hand_kp_tea"], g["hand_visible"], (64, 64), 2, (256, 256))
    for ci, (t, w) in enumerate(got):
        np.testing.assert_array_equal(t, g[f"hand_target_{ci}"])
        np.testing.assert_array_equal(w, g[f"hand_weight_{ci}"])
    for kind in ("Gaussian", "Cauchy"):
        for v in ("ori", "stu", "tea"):
            t, w = R.loader_labelmaps_animal(g[f"animal_{kind}_pts_{v}"], g[f"animal_{kind}_gate_{v}"], g[f"animal_{kind}_w0_{v}"], 64, 1.0, kind)
            np.testing.assert_array_equal(t.numpy(), g[f"animal_{kind}_target_{v}"])
            np.testing.assert_array_equal(w.numpy(), g[f"animal_{kind}_weight_{v}"])


def _split_like(flat, params):
    out, off = [], 0
    for p in params:
        out.append(T(flat[off:off + p.numel()]).view(p.shape).clone())
        off += p.numel()
    return out


def test_ema(golden):
    g = golden("ema")
    shapes = [(8, 3, 3, 3), (8,), (8,), (8,), (5, 8, 1, 1), (5,), (3, 7), (3,)]
    protos = [torch.empty(s) for s in shapes]
    student = _split_like(g["student0"], protos)
    teacher = _split_like(g["teacher0"], protos)
    R.ema_init(teacher, student)
    np.testing.assert_array_equal(np.concatenate([t.numpy().ravel() for t in teacher]), g["teacher_init"])
    for step in range(3):
        for s, d in zip(student, _split_like(g["deltas"][step], protos)):
            s.add_(d)
        R.ema_step(teacher, student, 0.999)
        np.testing.assert_array_equal(np.concatenate([t.numpy().ravel() for t in teacher]),
                                      g[f"teacher_step{step}"])
    ema = _split_like(g["mema_before"], protos)
    model = _split_like(g["mema_model"], protos)
    R.model_ema_update(ema, model, 0.99)
    np.testing.assert_array_equal(np.concatenate([t.numpy().ravel() for t in ema]), g["mema_after"])


OPTIM_CASES = {
    "adam": ("adam", dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)),
    "adamwd": ("adam", dict(lr=3e-4, betas=(0.8, 0.99), eps=1e-6, weight_decay=1e-2)),
    "sgd": ("sgd", dict(lr=0.1, momentum=0.9, dampening=0.0, weight_decay=0.0001, nesterov=True)),
    "sgdplain": ("sgd", dict(lr=0.05, momentum=0.8, dampening=0.1, weight_decay=0.0, nesterov=False)),
}
OPTIM_SHAPES = [(8, 3, 3, 3), (8,), (8,), (8,), (5, 8, 1, 1), (5,), (3, 7), (3,)]


@pytest.mark.parametrize("tag", sorted(OPTIM_CASES))
def test_student_teacher_step(golden, tag):
    """scaler.step(Adam | SGD) + OldWeightEMA.step + scaler.update() as run by torch on the CPU
    (train_human.py:436-440): the restatement must reproduce every intermediate state exactly."""
    g = golden("optim")
    algo, hyper = OPTIM_CASES[tag]
    protos = [torch.empty(s) for s in OPTIM_SHAPES]
    student = _split_like(g[f"{tag}_student0"], protos)
    teacher = [s.clone() for s in student]                      # OldWeightEMA.__init__ (utils.py:18-19)
    state1 = [torch.zeros_like(s) for s in student]
    state2 = [torch.zeros_like(s) for s in student]
    step, skipped = 0, []
    for it in range(5):
        grads = _split_like(g[f"{tag}_grads{it}"], protos)
        found_inf, step = R.student_teacher_step(algo, student, grads, state1, state2, teacher, step,
                                                 float(g[f"{tag}_scale{it}"]), 0.99, **hyper)
        skipped.append(found_inf)
        np.testing.assert_array_equal(np.concatenate([t.numpy().ravel() for t in student]), g[f"{tag}_student{it + 1}"])
        np.testing.assert_array_equal(np.concatenate([t.numpy().ravel() for t in teacher]), g[f"{tag}_teacher{it + 1}"])
    assert skipped == [False, False, True, False, False] and step == 4
    if algo == "adam":
        np.testing.assert_array_equal(np.concatenate([t.numpy().ravel() for t in state1]), g[f"{tag}_exp_avg"])
        np.testing.assert_array_equal(np.concatenate([t.numpy().ravel() for t in state2]), g[f"{tag}_exp_avg_sq"])
        assert float(g[f"{tag}_steps"]) == 4.0
    else:
        np.testing.assert_array_equal(np.concatenate([t.numpy().ravel() for t in state1]), g[f"{tag}_momentum_buffer"])


@pytest.mark.parametrize("tag", ["warp", "ragged", "stream", "mid"])
def test_style_loss(golden, tag):
    """adain/net.py:137-143 + autograd through calc_mean_std (the decoder pre-training job)."""
    g = golden("style_loss")
    x = T(g[f"{tag}_input"]).clone().requires_grad_(True)
    loss = R.calc_style_loss(x, T(g[f"{tag}_target"]))
    (gx,) = torch.autograd.grad(loss * 100.0, (x,))
    assert torch.equal(loss.detach(), T(g[f"{tag}_loss"]))
    assert torch.equal(gx, T(g[f"{tag}_grad"]))
    x2 = T(g[f"{tag}_input"]).clone().requires_grad_(True)
    m, s = R.calc_mean_std(x2)
    (g2,) = torch.autograd.grad([m, s], (x2,), [T(g[f"{tag}_dmean"]), T(g[f"{tag}_dstd"])])
    assert torch.equal(g2, T(g[f"{tag}_dfeat"]))


@pytest.mark.parametrize("tag", ["human", "animal"])
def test_channel_clamp(golden, tag):
    g = golden("clamp")
    y = R.channel_clamp(torch.from_numpy(g[f"{tag}_x"]), torch.from_numpy(g[f"{tag}_lo"]), torch.from_numpy(g[f"{tag}_hi"]))
    np.testing.assert_array_equal(y.contiguous().numpy(), g[f"{tag}_y"])


# ---- f1: affine re-warp loops (train_human.py:359-372, :385-412, :417-423) ------------------------
def aug_of(a):
    a = T(a)
    return [a[:, 0], [a[:, 1].long(), a[:, 2].long()], [a[:, 3], a[:, 4]], a[:, 5]]


@pytest.mark.parametrize("tag,k", [("tea1", 1), ("tea2", 2), ("teaodd", 1)])
def test_teacher_recon(golden, tag, k):
    g = golden("rewarp")
    views = [T(g[f"{tag}_in{i}"]) for i in range(k)]
    augs = [aug_of(g[f"{tag}_v{i}_aug"]) for i in range(k)]
    assert torch.equal(R.teacher_recon(views, augs, 4.0), T(g[f"{tag}_out"]))
    if k == 1:  # the composed-index restatement the CUDA kernel follows
        assert torch.equal(R.recon_restated(views[0], augs[0], 4.0), T(g[f"{tag}_out"]))


@pytest.mark.parametrize("tag,dt", [("stu16", torch.float16), ("stubf", torch.bfloat16), ("stu32", torch.float32)])
def test_student_recon(golden, tag, dt):
    g = golden("rewarp")
    y = T(g[f"{tag}_in"]).to(dt)
    aug = aug_of(g[f"{tag}_aug"])
    assert torch.equal(R.student_recon(y, aug, 4.0).float(), T(g[f"{tag}_out"]))
    # composition of the three source-index maps, first grid in the half dtype (autocast rule)
    ac = None if dt == torch.float32 else dt
    assert torch.equal(R.recon_restated(y, aug, 4.0, ac).float(), T(g[f"{tag}_out"]))
    # gradient = scatter-add along the composed map of G as it arrives in the student dtype (float32
    # sums, one rounding to dt; the reference rounds to dt after every call, hence the tolerance)
    G = T(g[f"{tag}_G"]).to(dt).float()
    b, c, h, w = y.shape
    grad = torch.zeros(b, c, h * w)
    angle, [tx, ty], [sx, sy], sc = aug
    for i in range(b):
        src = T(R.recon_source_index(angle[i].item(), tx[i].item(), ty[i].item(), sx[i].item(), sy[i].item(),
                                     sc[i].item(), 4.0, h, w, dt, ac))
        ok = src >= 0
        grad[i].index_add_(1, src[ok], G[i].reshape(c, -1)[:, ok])
    got = grad.reshape(b, c, h, w).to(dt).float()
    # bf16 (an extension: the trainers autocast to fp16): the reference's three intermediate bf16
    # roundings of sums of up to ~4 terms differ from one rounding by up to 2 ulp = 1.6 %
    tol = {torch.float32: 1e-5, torch.float16: 1e-2, torch.bfloat16: 2e-2}[dt]
    assert_close_scaled(got, T(g[f"{tag}_grad"]), tol, f"{tag} grad")


def test_affine_restatement_vs_torchvision():
    """oracle.affine_nearest_restated == torchvision tF.affine(nearest) bit for bit (float32 grids and
    half grids under autocast), random parameters, square / ragged sizes."""
    from torchvision.transforms import functional as tF

    rng = np.random.RandomState(3)
    for (h, w) in [(64, 64), (37, 53), (16, 128)]:
        img = torch.randn(2, h, w, generator=torch.Generator().manual_seed(h))
        for _ in range(25):
            ang, sc = float(rng.uniform(-180, 180)), float(rng.uniform(0.5, 1.7))
            sh = [float(rng.uniform(-30, 30)), float(rng.uniform(-30, 30))]
            tr = [float(rng.randint(-13, 14)) / 4, float(rng.randint(-13, 14)) / 4]
            assert torch.equal(tF.affine(img, ang, translate=tr, scale=sc, shear=sh),
                               R.affine_nearest_restated(img, ang, tr, sc, sh))
            for dt in (torch.float16, torch.bfloat16):
                with torch.autocast("cpu", dtype=dt):
                    ref = tF.affine(img.to(dt), ang, translate=tr, scale=sc, shear=sh)   # half image
                    ref32 = tF.affine(img, ang, translate=tr, scale=sc, shear=sh)       # float32 image, half bmm
                assert torch.equal(ref, R.affine_nearest_restated(img.to(dt), ang, tr, sc, sh, dt).float())
                # (torchvision rounds a float32 image through the half grid's dtype before sampling)
                assert torch.equal(ref32, R.affine_nearest_restated(img.to(dt).float(), ang, tr, sc, sh, dt))


def test_occlusion(golden):
    g = golden("rewarp")
    ratio, rate, size, image = g["occ_args"]
    rng = np.random.RandomState(int(g["occ_seed"]))
    out = R.occlude_keypoints(T(g["occ_in"]), T(g["occ_conf_table"]), g["occ_pred_position"], aug_of(g["occ_aug"]),
                              float(ratio), float(rate), int(size), int(image), rng=rng)
    assert torch.equal(out, T(g["occ_out"]))

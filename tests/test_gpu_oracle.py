"""GPU parity against the CPU oracle on seeded synthetic inputs at the BASELINE.json config
sizes (C1 hand 32x21, C2/C3 human 32x16, C4 animal 64x18 sigma=1.0, C5 microbench 256x21),
plus size-independent properties at full size.  All calls go through the C-ABI.
"""
import numpy as np
import pytest
import torch

import uda_poseestimation_b200 as U
from conftest import assert_close_scaled
from oracle import reference_port as R
from uda_poseestimation_b200 import synthetic as S

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- AdaIN ---------------------------
@pytest.mark.parametrize("n", [4, 32])
def test_adain_fp32_vs_oracle(dev, n):
    c, s = S.vgg_features(n, seed=1234)
    alpha = float(np.random.RandomState(1234).uniform(0, 1))
    ref = R.adain_mix(c, s, alpha)
    out = U.adain_mix(c.to(dev), s.to(dev), alpha)
    assert_close_scaled(out, ref, 1e-5, "adain_mix fp32")
    m_ref, s_ref = R.calc_mean_std(c)
    m, sd = U.calc_mean_std(c.to(dev))
    assert_close_scaled(m, m_ref, 1e-5, "mean")
    assert_close_scaled(sd, s_ref, 1e-5, "std")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_adain_16bit_vs_oracle(dev, dtype):
    c, s = S.vgg_features(4, seed=99)
    c16, s16 = c.to(dtype), s.to(dtype)
    ref = R.adain_mix(c16.float(), s16.float(), 0.6)  # fp32 math on the same quantised inputs
    out = U.adain_mix(c16.to(dev), s16.to(dev), 0.6)
    assert out.dtype == dtype
    assert_close_scaled(out.float(), ref, 1e-2, f"adain {dtype}")
    m, sd = U.calc_mean_std(c16.to(dev))
    m_ref, s_ref = R.calc_mean_std(c16.float())
    assert_close_scaled(m.float(), m_ref, 1e-2, "mean")
    assert_close_scaled(sd.float(), s_ref, 1e-2, "std")


@pytest.mark.parametrize("shape", [(2, 3, 1, 1), (1, 2, 7, 9), (2, 5, 64, 64), (1, 2, 100, 100), (1, 1, 256, 256),
                                   (3, 4, 8, 8), (2, 2, 16, 24)])
def test_adain_shapes(dev, shape):
    """tails, non-vectorisable planes, planes larger than a warp's registers, hw == 1 (NaN std)."""
    g = torch.Generator().manual_seed(5)
    c = torch.randn(*shape, generator=g)
    s = torch.randn(shape[0], shape[1], 11, 5, generator=g) * 3 + 1
    assert_close_scaled(U.adain_mix(c.to(dev), s.to(dev), 0.25), R.adain_mix(c, s, 0.25), 1e-5, f"adain {shape}")
    m, sd = U.calc_mean_std(c.to(dev))
    m_ref, s_ref = R.calc_mean_std(c)
    assert_close_scaled(m, m_ref, 1e-5, "mean")
    assert_close_scaled(sd, s_ref, 1e-5, "std")


def test_adain_properties_full_size(dev):
    """N=32 (67 MB per tensor): the output planes carry the style statistics; alpha=0 is the
    identity; alpha mixing is affine in alpha."""
    c, s = S.vgg_features(32, seed=7)
    c, s = c.to(dev), s.to(dev)
    t = U.adaptive_instance_normalization(c, s)
    m_t, s_t = U.calc_mean_std(t)
    m_s, s_s = U.calc_mean_std(s)
    # near-constant content planes: std(out)/std(style) = sqrt(var_c/(var_c+eps)) is visibly < 1
    const = U.calc_mean_std(c)[1].flatten() < 0.3
    assert_close_scaled(m_t.flatten()[~const], m_s.flatten()[~const], 1e-4, "mean(adain) == mean(style)")
    assert_close_scaled(s_t.flatten()[~const], s_s.flatten()[~const], 1e-3, "std(adain) == std(style)")
    assert torch.equal(U.adain_mix(c, s, 0.0), c)
    half = U.adain_mix(c, s, 0.5)
    assert_close_scaled(half, 0.5 * (t + c), 1e-5, "affine in alpha")


# ---------------------------------------------------------------- decode / PCK ---------------------
@pytest.mark.parametrize("cfg", ["C1", "C4", "C5"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_decode_vs_oracle(dev, cfg, dtype):
    b, k = S.CONFIGS[cfg]["batch"], S.CONFIGS[cfg]["joints"]
    hm = S.heatmaps(b, k, seed=1234).to(dtype)  # fp16 makes ties common
    preds_ref, max_ref = R.get_max_preds(hm.numpy())
    preds, maxvals = U.get_max_preds(hm.to(dev))
    np.testing.assert_array_equal(preds.cpu().numpy(), preds_ref)
    np.testing.assert_array_equal(maxvals.cpu().numpy(), max_ref)
    idx_ref = np.argmax(hm.numpy().reshape(b, k, -1), 2)
    np.testing.assert_array_equal(U.decode(hm.to(dev), want_idx=True)["idx"].cpu().numpy(), idx_ref)


def test_decode_bf16_and_adversarial(dev):
    adv = S.adversarial_heatmaps(k=4)
    for dtype in (torch.float32, torch.float16, torch.bfloat16):
        x = adv.to(dtype)
        ref_idx = torch.argmax(x.float().view(10, 4, -1), 2)
        ref_max = torch.amax(x.view(10, 4, -1), 2)
        r = U.decode(x.to(dev), want_idx=True, want_maxvals=True, want_preds=True)
        assert torch.equal(r["idx"].cpu().long(), ref_idx)
        np.testing.assert_array_equal(r["maxvals"].float().cpu().numpy()[..., 0], ref_max.float().numpy())
        p_ref, _ = R.get_max_preds_torch(x.float())
        assert torch.equal(r["preds"].cpu(), p_ref)


@pytest.mark.parametrize("cfg", ["C1", "C5"])
def test_accuracy_vs_oracle(dev, cfg):
    b, k = S.CONFIGS[cfg]["batch"], S.CONFIGS[cfg]["joints"]
    joints, vis = S.keypoints(b, k, seed=4321)
    target, _ = U.generate_target_batched(joints, vis, (64, 64), 2, (256, 256), device=dev)
    g = torch.Generator().manual_seed(1)
    pred = torch.roll(target.cpu(), shifts=(2, -3), dims=(2, 3)) + 0.02 * torch.randn(b, k, 64, 64, generator=g)
    for dtype in (torch.float32, torch.float16):  # train() feeds fp16 y_s, validate() fp32
        o = pred.to(dtype)
        acc_ref, avg_ref, cnt_ref, pred_ref = R.accuracy(o.numpy(), target.cpu().numpy())
        acc, avg, cnt, p = U.accuracy(o.to(dev), target)
        np.testing.assert_array_equal(acc, acc_ref)
        assert avg == avg_ref and cnt == cnt_ref
        np.testing.assert_array_equal(p, pred_ref)
        hits_ref, valid_ref, _ = R.pck_counts(o.numpy(), target.cpu().numpy())
        hits, valid, _ = U.pck_counts(o.to(dev), target)
        np.testing.assert_array_equal(hits.cpu().numpy(), hits_ref)
        np.testing.assert_array_equal(valid.cpu().numpy(), valid_ref)


def test_accuracy_adversarial(dev):
    adv = S.adversarial_heatmaps(k=4)
    other = torch.roll(adv, 1, dims=0)
    acc_ref, avg_ref, cnt_ref, pred_ref = R.accuracy(adv.numpy(), other.numpy())
    acc, avg, cnt, p = U.accuracy(adv.to(dev), other.to(dev))
    np.testing.assert_array_equal(acc, acc_ref)
    assert avg == avg_ref and cnt == cnt_ref
    np.testing.assert_array_equal(p, pred_ref)
    # no valid joint at all -> acc == -1 everywhere, avg 0, cnt 0
    z = torch.zeros(2, 3, 64, 64)
    acc, avg, cnt, _ = U.accuracy(z.to(dev), z.to(dev))
    assert (acc == -1).all() and avg == 0 and cnt == 0


# ---------------------------------------------------------------- losses ----------------------------
@pytest.mark.parametrize("cfg", ["C1", "C2", "C4"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_losses_vs_oracle(dev, cfg, dtype):
    b, k = S.CONFIGS[cfg]["batch"], S.CONFIGS[cfg]["joints"]
    rtol = 1e-5 if dtype == torch.float32 else 1e-2
    y_s = S.heatmaps(b, k, seed=10).to(dtype)
    label = S.heatmaps(b, k, seed=11, noise=0.0)
    weight = (torch.rand(b, k, 1, generator=torch.Generator().manual_seed(12)) > 0.1).float()
    scale = 65536.0  # GradScaler's initial scale (train_human.py:324,436)
    # oracle: autocast runs mse_loss / pow in fp32 on the (quantised) student output
    o_ref = y_s.float().clone().requires_grad_(True)
    l_ref = R.joints_mse_loss(o_ref, label, weight)
    (l_ref * scale).backward()
    o = y_s.to(dev).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16 if dtype != torch.bfloat16 else torch.bfloat16,
                        enabled=dtype != torch.float32):
        l = U.JointsMSELoss()(o, label.to(dev), weight.to(dev))
    assert l.dtype == torch.float32
    (l * scale).backward()
    assert o.grad.dtype == dtype
    assert_close_scaled(l.detach(), l_ref.detach(), 1e-5, "mse loss")  # fp32 accumulation either way
    assert_close_scaled(o.grad.float(), o_ref.grad, rtol, "mse grad")
    # ConsLoss against a rectified teacher with the k-th value mask
    tea = R.rectify(S.heatmaps(b, k, seed=13, peak=(0.3, 1.2)), S.CONFIGS[cfg]["sigma"])
    tea_mask = torch.rand(b, k, generator=torch.Generator().manual_seed(14)) > 0.5
    s_ref = y_s.float().clone().requires_grad_(True)
    c_ref = R.cons_loss(s_ref, tea, tea_mask=tea_mask)
    (c_ref * scale).backward()
    s = y_s.to(dev).requires_grad_(True)
    c = U.ConsLoss()(s, tea.to(dev), tea_mask=tea_mask.to(dev))
    (c * scale).backward()
    assert_close_scaled(c.detach().float(), c_ref.detach(), 1e-5, "cons loss")
    assert_close_scaled(s.grad.float(), s_ref.grad, rtol, "cons grad")


def test_loss_odd_shapes_and_reduction_none(dev):
    g = torch.Generator().manual_seed(3)
    for shape in [(2, 3, 5, 7), (1, 1, 1, 1), (3, 2, 9, 13), (2, 2, 8, 8)]:
        o = torch.randn(*shape, generator=g)
        t = torch.randn(*shape, generator=g)
        w = torch.rand(shape[0], shape[1], 1, generator=g)
        for red in ("mean", "none"):
            o1 = o.clone().requires_grad_(True)
            l1 = R.joints_mse_loss(o1, t, w, red)
            up = torch.rand(l1.shape, generator=g)
            l1.backward(up)
            o2 = o.to(dev).requires_grad_(True)
            l2 = U.JointsMSELoss(reduction=red)(o2, t.to(dev), w.to(dev))
            l2.backward(up.to(dev))
            assert_close_scaled(l2.detach(), l1.detach(), 1e-5, f"mse {red} {shape}")
            assert_close_scaled(o2.grad, o1.grad, 1e-5, f"mse grad {red} {shape}")
        vm = torch.rand(shape[0], shape[2], shape[3], generator=g) > 0.3
        tm = torch.rand(shape[0], shape[1], generator=g)  # float mask
        s1 = o.clone().requires_grad_(True)
        c1 = R.cons_loss(s1, t, valid_mask=vm, tea_mask=tm)
        c1.backward()
        s2 = o.to(dev).requires_grad_(True)
        c2 = U.ConsLoss()(s2, t.to(dev), valid_mask=vm.to(dev), tea_mask=tm.to(dev))
        c2.backward()
        assert_close_scaled(c2.detach(), c1.detach(), 1e-5, f"cons {shape}")
        assert_close_scaled(s2.grad, s1.grad, 1e-5, f"cons grad {shape}")


def test_adain_both_directions_in_one_launch(dev):
    """s2t and t2s (train_human.py:348-356) as ONE launch == two adain_mix launches, bit for bit, with a float and
    a device-resident alpha; and both against the oracle."""
    c1, s1 = S.vgg_features(3, seed=31, channels=64)
    c2, s2 = S.vgg_features(3, seed=32, channels=64)
    a2 = torch.tensor([0.8], device=dev)
    o1, o2 = U.adain_mix_multi([(c1.to(dev), s1.to(dev), 0.3), (c2.to(dev), s2.to(dev), a2)])
    assert torch.equal(o1, U.adain_mix(c1.to(dev), s1.to(dev), 0.3))
    assert torch.equal(o2, U.adain_mix(c2.to(dev), s2.to(dev), a2))
    assert_close_scaled(o1, R.adain_mix(c1, s1, 0.3), 1e-5, "s2t")
    assert_close_scaled(o2, R.adain_mix(c2, s2, 0.8), 1e-5, "t2s")
    # ragged planes take the generic path, one launch per job: same results
    r1, r2 = torch.rand(2, 5, 7, 9), torch.rand(2, 5, 3, 11)
    (g1,) = U.adain_mix_multi([(r1.to(dev), r2.to(dev), 0.5)])
    assert_close_scaled(g1, R.adain_mix(r1, r2, 0.5), 1e-5, "ragged")


# ---------------------------------------------------------------- masks / rectify --------------------
@pytest.mark.parametrize("cfg", ["C1", "C4"])
def test_teacher_targets_vs_oracle(dev, cfg):
    b, k, sigma = S.CONFIGS[cfg]["batch"], S.CONFIGS[cfg]["joints"], S.CONFIGS[cfg]["sigma"]
    hm = S.heatmaps(b, k, seed=77, peak=(0.3, 1.2))
    conf_ref, pos_ref, table_ref = R.confidence_mask(hm, 0.9)
    mask_ref, thresh_ref, act_ref = R.consistency_mask(hm, 0.5)
    rect_ref = R.rectify(hm, sigma)
    t = U.teacher_targets(hm.to(dev), sigma, 0.5, occlude_thresh=0.9)
    assert torch.equal(t["conf"].cpu(), conf_ref) and torch.equal(t["position"].cpu(), pos_ref)
    assert torch.equal(t["conf_table"].cpu(), table_ref)
    assert torch.equal(t["tea_mask"].cpu(), mask_ref)
    assert t["mask_thresh"].item() == np.float32(thresh_ref)
    rect = t["rectified"].cpu()
    assert torch.equal(rect != 0, rect_ref != 0) and torch.equal(rect == 1, rect_ref == 1)
    assert_close_scaled(rect, rect_ref, 1e-5, "rectified")
    assert_close_scaled(U.rectify(hm.to(dev), sigma), rect_ref, 1e-5, "rectify()")


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_conf_table_threshold_is_rounded_like_torch_for_half_maps(dev, dt):
    """`conf >= occlude_thresh` on a half tensor compares in half: torch rounds 0.9 to half(0.9) first, so a
    plane whose maximum IS half(0.9) passes (it fails against the float32 0.9).  Bit-exact bar."""
    b, k = 4, 5
    hm = (torch.rand(b, k, 16, 16) * 0.5).to(dt)
    edge = torch.tensor(0.9, dtype=dt)
    below = (edge.view(torch.int16) - 1).view(dt)      # the next representable value below half(0.9)
    hm[0, 0, 3, 3], hm[0, 1, 4, 4], hm[1, 2, 5, 5] = edge, below, 1.0
    conf_ref, pos_ref, table_ref = R.confidence_mask(hm, 0.9)
    assert bool(table_ref[0, 0]) and not bool(table_ref[0, 1]) and bool(table_ref[1, 2])
    conf, pos, table = U.confidence_mask(hm.to(dev), 0.9)
    assert torch.equal(table.cpu(), table_ref) and torch.equal(pos.cpu(), pos_ref)
    assert torch.equal(conf.cpu().to(dt), conf_ref)


def test_mask_select_all_ranks(dev):
    """every k of kthvalue on data with ties, NaN and infinities (bit-exact)"""
    g = torch.Generator().manual_seed(8)
    act = (torch.randn(97, generator=g) * 2).round() / 2
    act[5] = float("inf")
    act[6] = float("-inf")
    act[7] = -0.0
    for kth in range(1, 98):
        ratio = (kth + 0.5) / 97
        assert int(ratio * 97) == kth
        th_ref = torch.kthvalue(act, kth)[0]
        mask, th = U.consistency_mask(act.to(dev), ratio)
        assert th.cpu() == th_ref
        assert torch.equal(mask.cpu(), act > th_ref)
    with pytest.raises(IndexError):
        U.consistency_mask(act.to(dev), 0.0)
    # NaN sorts last (torch.kthvalue): the top rank is NaN and nothing is > NaN
    act[9] = float("nan")
    mask, th = U.consistency_mask(act.to(dev), 1.0)
    assert torch.isnan(th).item() and torch.isnan(torch.kthvalue(act, 97)[0]).item() and not mask.any()
    mask, th = U.consistency_mask(act.to(dev), 96.5 / 97)
    assert th.cpu() == torch.kthvalue(act, 96)[0] and torch.equal(mask.cpu(), act > th.cpu())


# ---------------------------------------------------------------- target writers ----------------------
@pytest.mark.parametrize("cfg,rectify", [("C1", True), ("C2", False), ("C4", True), ("C5", False), ("C5", True)])
def test_decode_select_equals_decode_then_select(dev, cfg, rectify):
    """udape_decode_select (k-th select in the last CTA of the decode launch; C5 without a rectified map takes
    the TMA-staged persistent kernel) against the two-launch sequence and the oracle: identical bits, also
    when the ticket word is reused by back-to-back launches and with a tea_mask_in / NaN planes."""
    from uda_poseestimation_b200.keypoint_detection import decode
    c = S.CONFIGS[cfg]
    b, k = c["batch"], c["joints"]
    hm = S.heatmaps(b, k, seed=33, peak=(0.3, 1.2))
    hm[0, 1] = float("nan")
    hm[1, 0] = -1.0
    tm = (torch.rand(b, k, generator=torch.Generator().manual_seed(3)) > 0.3).float()
    x = hm.to(dev)
    for ratio in (0.5, 1.0 / (b * k) + 1e-9, 1.0):
        kth = max(1, int(ratio * b * k))
        for tea_mask in (None, tm):
            two = decode(x, want_preds=True, want_maxvals_f32=True, rectify_sigma=2.0 if rectify else None)
            m2, t2 = U.consistency_mask(two["maxvals_f32"], kth / (b * k) + 1e-12, None if tea_mask is None else tea_mask.to(dev))
            for _ in range(3):
                one = decode(x, want_preds=True, want_maxvals_f32=True, rectify_sigma=2.0 if rectify else None,
                             select_kth=kth, select_tea_mask=None if tea_mask is None else tea_mask.to(dev))
                assert torch.equal(one["tea_mask"], m2)
                assert torch.equal(one["mask_thresh"].view(1).view(torch.int32), t2.view(1).view(torch.int32))
                assert torch.equal(one["preds"], two["preds"])
                assert torch.equal(one["maxvals_f32"].view(torch.int32), two["maxvals_f32"].view(torch.int32))
                if rectify:
                    assert torch.equal(one["rectified"].view(torch.int32), two["rectified"].view(torch.int32))
    # oracle: torch.kthvalue + the reference expression
    ref_mask, ref_thresh, _ = R.consistency_mask(hm, 0.5)
    one = decode(x, want_maxvals_f32=True, select_kth=int(0.5 * b * k))
    assert torch.equal(one["tea_mask"].cpu(), ref_mask)
    assert float(one["mask_thresh"]) == float(ref_thresh) or (np.isnan(float(ref_thresh)) and np.isnan(float(one["mask_thresh"])))


def test_generate_target_out_argument(dev):
    joints, vis = S.keypoints(4, 16, seed=8)
    t0, w0 = U.generate_target_batched(joints, vis, (64, 64), 2, (256, 256), device=dev)
    t1, w1 = torch.full_like(t0, 7.0), torch.full_like(w0, 7.0)
    r = U.generate_target_batched(joints, vis, (64, 64), 2, (256, 256), device=dev, out=(t1, w1))
    assert r[0] is t1 and torch.equal(t0, t1) and torch.equal(w0, w1)
    with pytest.raises(ValueError):
        U.generate_target_batched(joints, vis, (64, 64), 2, (256, 256), device=dev, out=(t1[:, :8], w1))


@pytest.mark.parametrize("cfg", ["C1", "C4"])
def test_generate_target_vs_oracle(dev, cfg):
    b, k, sigma = S.CONFIGS[cfg]["batch"], S.CONFIGS[cfg]["joints"], S.CONFIGS[cfg]["sigma"]
    joints, vis = S.keypoints(b, k, seed=31)
    ref = [R.generate_target(joints[i], vis[i], (64, 64), sigma, (256, 256)) for i in range(b)]
    ref_t, ref_w = np.stack([r[0] for r in ref]), np.stack([r[1] for r in ref])
    t, w = U.generate_target_batched(joints, vis, (64, 64), sigma, (256, 256), device=dev)
    np.testing.assert_array_equal(w.cpu().numpy(), ref_w)
    np.testing.assert_array_equal(t.cpu().numpy() != 0, ref_t != 0)
    assert_close_scaled(t, ref_t, 1e-5, "generate_target")
    # animal variant
    pts = torch.from_numpy(joints / 4.0).float()
    imgs, viss = [], []
    for i in range(min(b, 8)):
        for j in range(k):
            im, v = R.draw_labelmap_ori(torch.zeros(64, 64), pts[i, j], sigma)
            imgs.append(im.numpy())
            viss.append(v)
    img, v = U.draw_labelmap_batched(pts[:8], 64, 64, sigma)
    np.testing.assert_array_equal(v.cpu().numpy().ravel(), np.array(viss))
    assert_close_scaled(img.reshape(-1, 64, 64), np.stack(imgs), 1e-6, "draw_labelmap")


# ---------------------------------------------------------------- EMA -----------------------------------
def test_ema_pose_resnet101_bit_exact(dev):
    """The full 323-tensor / 52 992 853-parameter PoseResNet-101 census (K=21), 2 steps."""
    shapes = S.pose_resnet_param_shapes(21)
    student_cpu = S.parameter_list(shapes, seed=1)
    teacher_cpu = S.parameter_list(shapes, seed=2)

    class Bag(torch.nn.Module):
        def __init__(self, tensors):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in tensors])

    student, teacher = Bag(student_cpu).to(dev), Bag(teacher_cpu).to(dev)
    opt = U.OldWeightEMA(teacher, student, alpha=0.999)
    R.ema_init(teacher_cpu, student_cpu)
    for step in range(2):
        for t in student_cpu:
            t.mul_(1.01).add_(0.001 * (step + 1))
        with torch.no_grad():
            for p in student.parameters():
                p.mul_(1.01).add_(0.001 * (step + 1))
        opt.step()
        R.ema_step(teacher_cpu, student_cpu, 0.999)
    for p, ref in zip(teacher.parameters(), teacher_cpu):
        assert torch.equal(p.detach().cpu(), ref)


def test_ema_16bit_and_stale_plan(dev):
    torch.manual_seed(0)
    for dtype in (torch.bfloat16, torch.float16):
        a = torch.nn.Linear(33, 17).to(dev, dtype)
        b_ = torch.nn.Linear(33, 17).to(dev, dtype)
        ref_t = [p.detach().float().cpu().clone() for p in b_.parameters()]
        opt = U.OldWeightEMA(a, b_, alpha=0.9)
        with torch.no_grad():
            for p in b_.parameters():
                p.add_(1.0)
        src = [p.detach().float().cpu() for p in b_.parameters()]
        opt.step()
        for p, t0, s in zip(a.parameters(), ref_t, src):
            expect = (t0 * 0.9 + s * (1.0 - 0.9)).to(dtype)
            assert_close_scaled(p.detach().float(), expect.float(), 1e-2, f"ema {dtype}")
    # re-pointing a parameter's storage is detected and the plan is rebuilt
    a = torch.nn.Linear(8, 8).to(dev)
    b_ = torch.nn.Linear(8, 8).to(dev)
    opt = U.OldWeightEMA(a, b_, alpha=0.5)
    opt.step()
    with torch.no_grad():
        b_.weight.data = torch.full_like(b_.weight, 3.0)
    before = a.weight.detach().clone()
    opt.step()
    assert torch.equal(a.weight.detach(), before * 0.5 + 3.0 * 0.5)


@pytest.mark.parametrize("shape", [(4, 64, 256, 256), (4, 128, 128, 128), (4, 256, 64, 64), (4, 512, 32, 32)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_style_loss_pretrain_sizes_vs_oracle(dev, shape, dtype):
    """The four encoder levels of the decoder pre-training job (adain/net.py:158-161; relu1_1 .. relu4_1
    of a 256x256 crop): statistics, style loss and its gradient against the CPU oracle."""
    g = torch.Generator().manual_seed(shape[1])
    x = torch.relu(torch.randn(*shape, generator=g) + 0.2)
    t = torch.relu(torch.randn(*shape, generator=g) * 1.5 + 0.4)
    rtol = 1e-5 if dtype == torch.float32 else 1e-2
    xr = x.to(dtype).float().requires_grad_(True)
    tr = t.to(dtype).float()
    loss_ref = R.calc_style_loss(xr, tr)
    (g_ref,) = torch.autograd.grad(loss_ref, (xr,))
    xd = x.to(dtype).to(dev).requires_grad_(True)
    mean, std = U.calc_mean_std(xd.detach())
    m_ref, s_ref = R.calc_mean_std(xr.detach())
    assert_close_scaled(mean.float(), m_ref, rtol, "mean")
    assert_close_scaled(std.float(), s_ref, rtol, "std")
    loss = U.calc_style_loss(xd, t.to(dtype).to(dev))
    (gx,) = torch.autograd.grad(loss, (xd,))
    assert gx.dtype == dtype and gx.shape == xd.shape
    assert_close_scaled(loss.float(), loss_ref, 10 * rtol if dtype != torch.float32 else rtol, "style loss")
    if dtype == torch.float32:
        assert_close_scaled(gx, g_ref, rtol, "style-loss gradient")


def test_mean_std_backward_properties(dev):
    """Size-independent properties of the backward: the gradient of sum(mean) is 1/hw everywhere, the
    gradient of a plane's std is orthogonal to constants (sums to zero) and has norm 1/sqrt(hw-1)
    (up to the eps term), and planes do not leak into each other."""
    torch.manual_seed(1)
    x = (torch.randn(3, 5, 48, 48, device=dev) * 2 + 1).requires_grad_(True)
    mean, std = U.calc_mean_std(x)
    (g,) = torch.autograd.grad(mean.sum(), (x,), retain_graph=True)
    assert torch.equal(g, torch.full_like(g, 1.0 / (48 * 48)))
    w = torch.zeros_like(std)
    w[1, 2] = 1.0
    (g,) = torch.autograd.grad((std * w).sum(), (x,))
    assert g[0].abs().max() == 0 and g[2].abs().max() == 0 and g[1, :2].abs().max() == 0
    plane = g[1, 2].double()
    assert abs(plane.sum().item()) < 1e-6
    var = x[1, 2].double().var().item()
    expect = (var / (var + 1e-5)) ** 0.5 / (48 * 48 - 1) ** 0.5
    assert abs(plane.norm().item() - expect) < 1e-6 * expect + 1e-9


def test_style_transfer_forward_only_vs_oracle(dev):
    """StyleTransfer = what the trainers keep of Style_net.Net.forward (element [2]) + the clamp: a small
    encoder with the reference's 31-layer slicing and a mirrored decoder, cuDNN convolutions in fp32."""
    nn = torch.nn

    def make():
        torch.manual_seed(7)
        layers, c = [nn.Conv2d(3, 3, 1)], 3
        for i, width in enumerate((8, 8, 12, 12, 16, 16, 16, 16, 16, 24)):   # 1 + 10 * 3 = 31 children
            layers += [nn.ReflectionPad2d(1), nn.Conv2d(c, width, 3), nn.ReLU()]
            c = width
        enc = nn.Sequential(*layers, nn.Conv2d(c, c, 1))                     # a 32nd child the slicing must ignore
        dec = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(24, 8, 3), nn.ReLU(), nn.Conv2d(8, 3, 1))
        return enc, dec

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        enc, dec = make()
        g = torch.Generator().manual_seed(2)
        content = torch.randn(3, 3, 40, 40, generator=g)
        style = torch.randn(3, 3, 40, 40, generator=g) * 1.5 + 0.3
        lo, hi = torch.tensor([-0.4, -0.5, -0.3]), torch.tensor([0.5, 0.4, 0.6])
        enc_d, dec_d = make()
        st = U.StyleTransfer(enc_d, dec_d).to(dev).eval()
        assert all(not p.requires_grad for n_, p in st.named_parameters() if n_.startswith("enc_"))
        with torch.no_grad():
            out = st(content.to(dev), style.to(dev), 0.6)
            assert out[0] is None and out[1] is None
            assert_close_scaled(out[2], R.style_transfer(enc, dec, content, style, 0.6), 1e-4, "g_t")
            y = st.stylize(content.to(dev), style.to(dev), 0.6, lo.to(dev), hi.to(dev))
            assert y.is_contiguous()
            assert_close_scaled(y, R.style_transfer(enc, dec, content, style, 0.6, lo, hi).contiguous(), 1e-4, "stylize")
            # different spatial sizes for content and style take the two-pass encode
            style2 = torch.randn(3, 3, 24, 56, generator=g)
            assert_close_scaled(st(content.to(dev), style2.to(dev), 1.0)[2], R.style_transfer(enc, dec, content, style2, 1.0),
                                1e-4, "g_t (ragged style)")
    finally:
        torch.backends.cudnn.allow_tf32 = old


# ---------------------------------------------------------------- clamp + fused loss step ------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(32, 3, 256, 256), (2, 3, 7, 9), (1, 5, 33, 1)])
def test_channel_clamp_vs_oracle(dev, dtype, shape):
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(*shape, generator=g) * 2).to(dtype)
    c = shape[1]
    lo = -torch.rand(c, generator=g) - 0.3
    hi = torch.rand(c, generator=g) + 0.3
    ref = R.channel_clamp(x.float(), lo, hi).to(dtype)  # bounds are fp32; selection commutes with rounding of x only
    # values selected from the bounds are rounded to x's dtype by the store
    y = U.channel_clamp(x.to(dev), lo.to(dev), hi.to(dev))
    assert y.is_contiguous() and y.dtype == dtype
    assert torch.equal(y.cpu(), ref.contiguous())
    # in place
    xd = x.to(dev)
    U.channel_clamp(xd, lo.to(dev), hi.to(dev), out=xd)
    assert torch.equal(xd.cpu(), ref.contiguous())


@pytest.mark.parametrize("cfg", ["C1", "C2", "C4"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_fused_losses_vs_oracle(dev, cfg, dtype):
    """udape_loss_step == JointsMSELoss + lambda_c*ConsLoss of the reference, and the gradients of
    scale*loss_all from autograd through the reference modules (train_human.py:425-436)."""
    b, k, sigma = S.CONFIGS[cfg]["batch"], S.CONFIGS[cfg]["joints"], S.CONFIGS[cfg]["sigma"]
    rtol = 1e-5 if dtype == torch.float32 else 1e-2
    scale, lam = 65536.0, 0.7
    y_s = S.heatmaps(b, k, seed=20).to(dtype)
    y_t = S.heatmaps(b, k, seed=21).to(dtype)
    label = S.heatmaps(b, k, seed=22, noise=0.0)
    weight = (torch.rand(b, k, 1, generator=torch.Generator().manual_seed(23)) > 0.1).float()
    tea_raw = S.heatmaps(b, k, seed=24, peak=(0.3, 1.2))
    tea_raw[0, 0] = -1.0  # all-negative plane: rectify pastes at (0, 0)
    tea = R.rectify(tea_raw, sigma)
    tea_mask, _, _ = R.consistency_mask(tea_raw, 0.5)
    o1 = y_s.float().clone().requires_grad_(True)
    o2 = y_t.float().clone().requires_grad_(True)
    l_s = R.joints_mse_loss(o1, label, weight)
    l_c = R.cons_loss(o2, tea, tea_mask=tea_mask)
    l_all = l_s + lam * l_c
    (l_all * scale).backward()
    gs = torch.full((1,), scale, device=dev)
    for route in ("materialised", "analytic"):
        if route == "materialised":
            losses, g1, g2 = U.fused_losses(y_s.to(dev), label.to(dev), weight.to(dev), y_t.to(dev), tea.to(dev),
                                            tea_mask.to(dev), lambda_c=lam, grad_scale=gs)
        else:
            d = U.decode(tea_raw.to(dev), want_preds=True)
            losses, g1, g2 = U.fused_losses(y_s.to(dev), label.to(dev), weight.to(dev), y_t.to(dev), None,
                                            tea_mask.to(dev), lambda_c=lam, grad_scale=scale,
                                            tea_preds=d["preds"], sigma=sigma)
        assert g1.dtype == dtype and g2.dtype == dtype
        assert_close_scaled(losses[0], l_all.detach(), 1e-5, f"{route} loss_all")
        assert_close_scaled(losses[1], l_s.detach(), 1e-5, f"{route} loss_s")
        assert_close_scaled(losses[2], l_c.detach(), 1e-5, f"{route} loss_c")
        assert_close_scaled(g1.float(), o1.grad, rtol, f"{route} grad y_s")
        assert_close_scaled(g2.float(), o2.grad, rtol, f"{route} grad y_t_stu")
    # the two routes use the same device functions for the rectified values: identical gradient bits
    # (the loss sums run over 8-element instead of 4-element groups on the analytic route: same
    # addends, different fp32 summation order)
    la, ga1, ga2 = U.fused_losses(y_s.to(dev), label.to(dev), weight.to(dev), y_t.to(dev), U.rectify(tea_raw.to(dev), sigma),
                                  tea_mask.to(dev), lambda_c=lam, grad_scale=scale)
    assert torch.equal(ga2, g2) and torch.equal(ga1, g1)
    assert_close_scaled(la, losses, 1e-6, "analytic vs materialised losses")


def test_fused_losses_partial_and_odd_shapes(dev):
    g = torch.Generator().manual_seed(9)
    for shape in [(2, 3, 5, 7), (1, 1, 1, 1), (3, 2, 9, 13)]:
        o = torch.randn(*shape, generator=g)
        t = torch.randn(*shape, generator=g)
        w = torch.rand(shape[0], shape[1], 1, generator=g)
        tm = torch.rand(shape[0], shape[1], generator=g) > 0.4
        o1 = o.clone().requires_grad_(True)
        l1 = R.joints_mse_loss(o1, t, w)
        l1.backward()
        losses, g1, g2 = U.fused_losses(o.to(dev), t.to(dev), w.to(dev), None)
        assert g2 is None
        assert_close_scaled(losses[1], l1.detach(), 1e-5, f"mse only {shape}")
        assert_close_scaled(g1, o1.grad, 1e-5, f"mse-only grad {shape}")
        assert float(losses[2]) == 0.0
        s1 = o.clone().requires_grad_(True)
        c1 = R.cons_loss(s1, t, tea_mask=tm)
        (2.0 * c1).backward()
        losses, g1, g2 = U.fused_losses(None, None, None, o.to(dev), t.to(dev), tm.to(dev), lambda_c=2.0)
        assert g1 is None
        assert_close_scaled(losses[2], c1.detach(), 1e-5, f"cons only {shape}")
        assert_close_scaled(losses[0], 2.0 * c1.detach(), 1e-5, f"cons only all {shape}")
        assert_close_scaled(g2, s1.grad, 1e-5, f"cons-only grad {shape}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(256, 21, 64, 64), (32, 16, 64, 64), (3, 5, 24, 24), (2, 2, 8, 8)])   # square: the reference's rectify raises on others (utils.py:89,101-105)
def test_fused_losses_pair_route_equals_plane_per_cta(dev, dtype, shape):
    """The analytic fused step has two grids: a plane per CTA, and — the default from 16 x SMs plane pairs up (C5) —
    one supervised + one consistency plane per CTA with the first loads of both issued up front.  Same per-thread
    order of additions, same block trees, same partial slots: losses and gradients agree bit for bit; and against
    the oracle at the usual bars."""
    import os
    b, k, h, w = shape
    y_s = S.heatmaps(b, k, seed=30, h=h, w=w).to(dtype)
    y_t = S.heatmaps(b, k, seed=31, h=h, w=w).to(dtype)
    tea_raw = S.heatmaps(b, k, seed=32, peak=(0.3, 1.2), h=h, w=w)
    label = S.heatmaps(b, k, seed=33, noise=0.0, h=h, w=w)
    weight = (torch.rand(b, k, 1, generator=torch.Generator().manual_seed(34)) > 0.1).float()
    tea_mask, _, _ = R.consistency_mask(tea_raw, 0.5)
    d = U.decode(tea_raw.to(dev), want_preds=True)
    args = (y_s.to(dev), label.to(dev), weight.to(dev), y_t.to(dev), None, tea_mask.to(dev))
    outs = {}
    for flag in ("0", "1"):
        outs[flag] = _with_env("UDAPE_LOSS_PAIR", flag, lambda: U.fused_losses(*args, lambda_c=0.7, grad_scale=65536.0,
                                                                              tea_preds=d["preds"], sigma=2))
    for a, e in zip(outs["1"], outs["0"]):
        assert torch.equal(a, e)
    default = U.fused_losses(*args, lambda_c=0.7, grad_scale=65536.0, tea_preds=d["preds"], sigma=2)
    for a, e in zip(default, outs["0"]):
        assert torch.equal(a, e)
    if b * k <= 1024:     # (the oracle's rectify loops over every plane on the host)
        o1, o2 = y_s.float().clone().requires_grad_(True), y_t.float().clone().requires_grad_(True)
        l_all = R.joints_mse_loss(o1, label, weight) + 0.7 * R.cons_loss(o2, R.rectify(tea_raw, 2), tea_mask=tea_mask)
        (l_all * 65536.0).backward()
        rtol = 1e-5 if dtype == torch.float32 else 1e-2
        assert_close_scaled(outs["1"][0][0], l_all.detach(), 1e-5, "pair route loss_all")
        assert_close_scaled(outs["1"][1].float(), o1.grad, rtol, "pair route grad y_s")
        assert_close_scaled(outs["1"][2].float(), o2.grad, rtol, "pair route grad y_t_stu")


# ---------------------------------------------------------------- TMA-staged vs register-staged paths --------
def _with_env(name, value, fn):
    import os
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_decode_and_pck_tma_equals_generic(dev, dtype):
    """>= 2048 planes route decode / PCK through the cp.async.bulk pipeline (csrc/pipeline.cuh); both
    routes must agree bit for bit with each other and with numpy/torch on adversarial planes (ties,
    -0.0/+0.0, all-negative, NaN/Inf, fp16 quantisation)."""
    adv = S.adversarial_heatmaps(k=4)                                    # [10,4,64,64]
    rnd = S.heatmaps(60, 21, seed=77)                                    # 1260 planes
    x = torch.cat([adv.reshape(-1, 64, 64), rnd.reshape(-1, 64, 64)] * 2).reshape(-1, 20, 64, 64).to(dtype)
    assert x.shape[0] * x.shape[1] >= 2048
    xd = x.to(dev)
    kw = dict(want_idx=True, want_maxvals=True, want_preds=True, want_position=True, occlude_thresh=0.9)
    tma = _with_env("UDAPE_NO_TMA", "0", lambda: U.decode(xd, **kw))
    gen = _with_env("UDAPE_NO_TMA", "1", lambda: U.decode(xd, **kw))
    b, k = x.shape[:2]
    ref_idx = torch.argmax(x.float().view(b, k, -1), 2)
    for r in (tma, gen):
        assert torch.equal(r["idx"].cpu().long(), ref_idx)
        assert torch.equal(r["preds"].cpu(), R.get_max_preds_torch(x.float())[0])
        np.testing.assert_array_equal(r["maxvals"].float().cpu().numpy()[..., 0],
                                      torch.amax(x.view(b, k, -1), 2).float().numpy())
    for key in tma:
        a, g = tma[key], gen[key]
        assert torch.equal(a.view(torch.uint8) if a.dtype == torch.bool else a.float().nan_to_num(nan=12345.0),
                           g.view(torch.uint8) if g.dtype == torch.bool else g.float().nan_to_num(nan=12345.0)), key
    other = torch.roll(x, 3, dims=0).float()
    ref = R.pck_counts(x.float().numpy() if dtype != torch.float16 else x.numpy(), other.numpy())
    for flag in ("0", "1"):
        hits, valid, pred = _with_env("UDAPE_NO_TMA", flag, lambda: U.pck_counts(xd, other.to(dev)))
        np.testing.assert_array_equal(hits.cpu().numpy(), ref[0])
        np.testing.assert_array_equal(valid.cpu().numpy(), ref[1])
        np.testing.assert_array_equal(pred.cpu().numpy(), ref[2])


def test_tickets_are_self_resetting(dev):
    """Back-to-back reductions reuse ticket words without any memset (include/udape.h "tickets")."""
    o = S.heatmaps(8, 16, seed=1).to(dev)
    t = S.heatmaps(8, 16, seed=2, noise=0.0).to(dev)
    first = [float(U.JointsMSELoss()(o, t)) for _ in range(3)]
    many = [float(U.JointsMSELoss()(o, t)) for _ in range(5000)]  # wraps the 4096-slot rotation
    assert len(set(first + many)) == 1
    from uda_poseestimation_b200 import _lib
    pool = _lib._ticket_pools[dev.index or 0]
    torch.cuda.synchronize()
    assert int(pool.buf.abs().sum()) == 0

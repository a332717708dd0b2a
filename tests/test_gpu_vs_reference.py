"""GPU parity against the REFERENCE ITSELF: the CUDA operators, through the C-ABI, beside the reference's own
functions (``oracle/reference_live.py``: imported from ``/root/reference`` in the build container, from the
bytecode ``oracle/build_ref.py`` compiled into ``oracle/_ref/`` on the GPU box — the directory travels with the
repository like the built ``.so``).  ``tests/test_gpu_oracle.py`` makes the same comparisons against the restated
port at every BASELINE config; this file removes the restatement from the loop for each ★ callable of SURVEY.md §8a.
Skipped where neither the tree nor the bytecode exists.
"""
import numpy as np
import pytest
import torch

import uda_poseestimation_b200 as U
from conftest import assert_close_scaled
from oracle import ref_loader
from oracle import reference_live as RL
from uda_poseestimation_b200 import synthetic as S

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="no reference tree / oracle/_ref bytecode")]


def test_where_the_reference_comes_from():
    assert ref_loader.kind() in ("source", "bytecode")
    fn = ref_loader.load("function")
    assert fn.calc_mean_std.__module__ == "_udape_ref_function"        # not the port's restatement
    assert RL.rectify.__module__ == "oracle.reference_live" and RL.teacher_recon.__module__ == "oracle.reference_port"


# a1-a3 ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [4, 32])
def test_adain_statistics_mix(dev, n):
    fn, sn = ref_loader.load("function"), ref_loader.load("style_net")
    c, s = S.vgg_features(n, seed=77)
    m, sd = U.calc_mean_std(c.to(dev))
    m_ref, s_ref = fn.calc_mean_std(c)
    assert_close_scaled(m, m_ref, 1e-5, "mean")
    assert_close_scaled(sd, s_ref, 1e-5, "std")
    assert_close_scaled(U.adaptive_instance_normalization(c.to(dev), s.to(dev)), fn.adaptive_instance_normalization(c, s), 1e-5, "adain")
    assert_close_scaled(U.adain(c.to(dev), s.to(dev)), sn.adain(c, s), 1e-5, "Style_net.adain")
    for alpha in (0.0, 0.37, 1.0):
        t = sn.adain(c, s)
        assert_close_scaled(U.adain_mix(c.to(dev), s.to(dev), alpha), alpha * t + (1 - alpha) * c, 1e-5, f"mix {alpha}")   # Style_net.py:167-168


# a4 / a5 ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sigma,size", [(2, (64, 64)), (1.0, (64, 64)), (2, (8, 8))])
def test_generate_target(dev, sigma, size):
    du = ref_loader.load("dataset_util")
    joints, vis = S.keypoints(16, 21, seed=5)
    tgt, wgt = U.generate_target_batched(joints, vis, size, sigma, (256, 256), device=dev)
    ref = [du.generate_target(joints[i], vis[i], size, sigma, (256, 256)) for i in range(16)]
    ref_t, ref_w = np.stack([r[0] for r in ref]), np.stack([r[1] for r in ref])
    np.testing.assert_array_equal(wgt.cpu().numpy(), ref_w)
    np.testing.assert_array_equal(tgt.cpu().numpy() != 0, ref_t != 0)          # integer placement: exact
    assert_close_scaled(tgt, ref_t, 1e-5, "generate_target")
    one_t, one_w = U.generate_target(joints[3], vis[3], size, sigma, (256, 256))
    np.testing.assert_array_equal(one_w, ref_w[3])
    assert_close_scaled(one_t, ref_t[3], 1e-5, "generate_target (single)")


@pytest.mark.parametrize("kind", ["Gaussian", "Cauchy"])
def test_draw_labelmap_ori(dev, kind):
    du = ref_loader.load("dataset_util")
    g = torch.Generator().manual_seed(9)
    pts = torch.cat([torch.rand(40, 2, generator=g) * 70 - 3, torch.tensor([[0.0, 0.0], [63.0, 63.0], [3.0, 3.0], [60.9, 3.2]])])
    for p in pts:
        img, v = U.draw_labelmap_ori(torch.zeros(64, 64, device=dev), p, 1.0, type=kind)
        img_ref, v_ref = du.draw_labelmap_ori(torch.zeros(64, 64), p, 1.0, type=kind)
        assert v == v_ref
        assert_close_scaled(img, torch.as_tensor(img_ref), 1e-5, f"labelmap {kind} {p.tolist()}")


# a6-a8, a11 ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["C2", "C4"])
def test_decode_rectify_accuracy(dev, cfg):
    kd, ut = ref_loader.load("keypoint_detection"), ref_loader.load("utils")
    b, k, sigma = S.CONFIGS[cfg]["batch"], S.CONFIGS[cfg]["joints"], S.CONFIGS[cfg]["sigma"]
    hm = S.heatmaps(b, k, seed=31, peak=(0.2, 1.2))
    hm[0, 0], hm[1, 1] = 0.0, -1.0
    for x in (hm, hm.half()):
        p_ref, m_ref = kd.get_max_preds(x.numpy())
        p, m = U.get_max_preds(x.to(dev))
        np.testing.assert_array_equal(p.cpu().numpy(), p_ref)
        np.testing.assert_array_equal(m.cpu().numpy(), m_ref)
    pt, mt = U.get_max_preds_torch(hm.to(dev))
    pt_ref, mt_ref = ut.get_max_preds_torch(hm)
    assert torch.equal(pt.cpu(), pt_ref) and torch.equal(mt.cpu(), mt_ref)
    rect_ref = ut.rectify(hm.clone(), sigma)
    rect = U.rectify(hm.to(dev), sigma)
    assert torch.equal(rect.cpu() != 0, rect_ref != 0) and torch.equal(rect.cpu() == 1, rect_ref == 1)
    assert_close_scaled(rect, rect_ref, 1e-5, "rectify")
    tgt = S.heatmaps(b, k, seed=32)
    for x in (hm, hm.half()):
        acc_ref, avg_ref, cnt_ref, pred_ref = kd.accuracy(x.numpy(), tgt.numpy())
        acc, avg, cnt, pred = U.accuracy(x.to(dev), tgt.to(dev))
        np.testing.assert_array_equal(acc, acc_ref)
        assert avg == avg_ref and cnt == cnt_ref
        np.testing.assert_array_equal(pred, pred_ref)
        acc_n, avg_n, cnt_n, pred_n = U.accuracy(x.numpy(), tgt.numpy())      # the reference's numpy signature
        np.testing.assert_array_equal(acc_n, acc_ref)
        assert avg_n == avg_ref and cnt_n == cnt_ref


# a9 / a10 -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_losses_forward_backward(dev, dtype):
    lo = ref_loader.load("loss")
    b, k = 32, 16
    y = S.heatmaps(b, k, seed=41).to(dtype)
    joints, vis = S.keypoints(b, k, seed=42)
    label, weight = U.generate_target_batched(joints, vis, (64, 64), 2, (256, 256), device=dev)
    tea = S.heatmaps(b, k, seed=43, peak=(0.3, 1.2))
    mask = torch.rand(b, k, generator=torch.Generator().manual_seed(44)) > 0.5
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for red in ("mean", "none"):
        out = U.JointsMSELoss(reduction=red)(y.to(dev), label, weight)
        ref = lo.JointsMSELoss(reduction=red)(y.float(), label.cpu(), weight.cpu())
        assert_close_scaled(out.float(), ref, 1e-5 if dtype == torch.float32 else 2e-3, f"JointsMSELoss {red}")
    o = y.to(dev).requires_grad_(True)
    o_ref = y.float().requires_grad_(True)
    loss = U.JointsMSELoss()(o, label, weight) + U.ConsLoss()(o, tea.to(dev), tea_mask=mask.to(dev))
    loss_ref = lo.JointsMSELoss()(o_ref, label.cpu(), weight.cpu()) + lo.ConsLoss()(o_ref, tea, tea_mask=mask)
    (loss * 65536.0).backward()
    (loss_ref * 65536.0).backward()
    assert_close_scaled(loss.float(), loss_ref, 1e-5 if dtype == torch.float32 else 2e-3, "loss")
    assert_close_scaled(o.grad.float(), o_ref.grad, tol, "grad")
    vm = torch.rand(b, 64, 64, generator=torch.Generator().manual_seed(45)) > 0.3
    out = U.ConsLoss()(y.to(dev), tea.to(dev), valid_mask=vm.to(dev), tea_mask=mask.to(dev))
    ref = lo.ConsLoss()(y.float(), tea, valid_mask=vm, tea_mask=mask)
    assert_close_scaled(out.float(), ref, 1e-5 if dtype == torch.float32 else 2e-3, "ConsLoss valid_mask")


# a12 / a13 + the optimizer tail -----------------------------------------------------------------------
def test_old_weight_ema_and_trainer_tail_bit_exact(dev):
    """OldWeightEMA.step() (utils.py:21-25) and scaler.step(Adam) + EMA (train_human.py:436-438): the reference's own
    objects on the CPU against the fused CUDA step — the EMA bit for bit, Adam to 1e-6 (torch's CPU Adam and the
    kernel both follow the single-tensor op order; FMA contraction differs)."""
    ut = ref_loader.load("utils")
    torch.manual_seed(3)
    mk = lambda: torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5))
    tea_c, stu_c = mk(), mk()
    tea_g, stu_g = mk().to(dev), mk().to(dev)
    stu_g.load_state_dict(stu_c.state_dict())
    ema_c, ema_g = ut.OldWeightEMA(tea_c, stu_c, alpha=0.97), U.OldWeightEMA(tea_g, stu_g, alpha=0.97)
    for pc, pg in zip(tea_c.parameters(), tea_g.parameters()):
        assert torch.equal(pc, pg.cpu())                     # the constructor's copy, utils.py:18-19
    with torch.no_grad():
        for pc, pg in zip(stu_c.parameters(), stu_g.parameters()):
            d = torch.randn(pc.shape)
            pc.add_(d)
            pg.add_(d.to(dev))
    for _ in range(3):
        ema_c.step()
        ema_g.step()
    for pc, pg in zip(tea_c.parameters(), tea_g.parameters()):
        assert torch.equal(pc.detach(), pg.detach().cpu()), "OldWeightEMA.step"
    # the tail
    shapes = [(64, 3, 7, 7), (64,), (19, 64), (19,)]
    st = [torch.randn(*s) * 0.05 for s in shapes]
    tail = RL.TrainerTail(st, [x.clone() for x in st], 1e-3, 0.999, 65536.0, "cpu")
    stu = torch.nn.ParameterList([torch.nn.Parameter(x.clone().to(dev)) for x in st])
    tea = torch.nn.ParameterList([torch.nn.Parameter(x.clone().to(dev), requires_grad=False) for x in st])
    ema = U.OldWeightEMA(tea, stu, alpha=0.999)
    opt = U.Adam(stu.parameters(), lr=1e-3)
    opt.attach_teacher(ema)
    g = torch.Generator().manual_seed(8)
    for it in range(4):
        grads = [torch.randn(*s, generator=g) * 65536.0 for s in shapes]
        if it == 2:
            grads[1][5] = float("inf")      # GradScaler skips the step (and halves the scale); the EMA still runs
        scale = float(tail.scaler.get_scale())
        dummy = torch.zeros((), requires_grad=True)
        tail.scaler.scale(dummy * 1.0).backward()
        tail.step(grads)
        for p, gr in zip(stu.parameters(), grads):
            p.grad = gr.to(dev)
        opt.grad_scale, opt.found_inf = torch.full((), scale, device=dev), opt.check_grads()
        opt.step()
        ema.step()
        for a, e in zip(list(stu.parameters()) + list(tea.parameters()), list(tail.student.parameters()) + list(tail.teacher.parameters())):
            assert_close_scaled(a.detach(), e.detach(), 2e-6, f"trainer tail, iteration {it}")


def test_model_ema(dev):
    """ModelEMA.update / momentum_update (lib/models/ema.py:18-44): the reference's class runs on CUDA tensors of the
    same GPU (its constructor moves the copy there, :9), the package's on the multi-tensor kernels."""
    em = ref_loader.load("ema")
    torch.manual_seed(5)
    mk = lambda: torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4)).to(dev)
    net_r, net_g = mk(), mk()
    net_g.load_state_dict(net_r.state_dict())
    ref, out = em.ModelEMA(net_r, 0.9), U.ModelEMA(net_g, 0.9)
    with torch.no_grad():
        for pr, pg in zip(net_r.parameters(), net_g.parameters()):
            d = torch.randn(pr.shape, device=dev)
            pr.add_(d)
            pg.add_(d)
        net_r[1].running_mean.add_(0.5)
        net_g[1].running_mean.add_(0.5)
    for _ in range(2):
        ref.update(net_r)
        out.update(net_g)
    for (n1, a), (n2, e) in zip(out.ema.state_dict().items(), ref.ema.state_dict().items()):
        assert n1 == n2
        assert_close_scaled(a.float(), e.float(), 1e-6, n1)
    ref.momentum_update(net_r, 0.8)
    out.momentum_update(net_g, 0.8)
    for a, e in zip(out.ema.parameters(), ref.ema.parameters()):
        assert_close_scaled(a, e, 1e-6, "momentum_update")

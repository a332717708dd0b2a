"""The assembled hot-path step (uda_poseestimation_b200.hotpath) against the CPU oracle running
the reference's sequence, eagerly (serial and forked streams) and as a replayed CUDA graph with
device-resident alpha scalars."""
import numpy as np
import pytest
import torch

import uda_poseestimation_b200 as U
from conftest import assert_close_scaled
from oracle import reference_port as R
from uda_poseestimation_b200 import rewarp as RW
from uda_poseestimation_b200 import synthetic as S
from uda_poseestimation_b200.hotpath import HotPathStep, StepInputs, step_algorithmic_bytes

pytestmark = pytest.mark.gpu


class Bag(torch.nn.Module):
    def __init__(self, tensors):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in tensors])


def _inputs(dev, b=4, k=16, n_feat_c=32, seed=5, rewarp=True):
    src, tgt_style = S.vgg_features(b, seed, channels=n_feat_c)
    tgt, src_style = S.vgg_features(b, seed + 1, channels=n_feat_c)
    joints, vis = S.keypoints(b, k, seed + 5)
    lab = [R.generate_target(joints[i], vis[i], (64, 64), 2, (256, 256)) for i in range(b)]
    host = dict(feat_src=src, feat_tgt_ori=tgt_style, feat_tgt_tea=tgt, feat_src_ori=src_style,
                y_s=S.heatmaps(b, k, seed + 2).half(), y_t_stu=S.heatmaps(b, k, seed + 3).half(),
                y_t_tea=S.heatmaps(b, k, seed + 4, peak=(0.3, 1.2)),
                label_s=torch.from_numpy(np.stack([x[0] for x in lab])),
                weight_s=torch.from_numpy(np.stack([x[1] for x in lab])))
    inp = StepInputs(**{n: t.to(dev) for n, t in host.items()}, alpha_s2t=torch.tensor([0.3], device=dev),
                     alpha_t2s=torch.tensor([0.8], device=dev))
    if rewarp:
        # the collated meta['aug_param_tea'] / ['aug_param_stu'] of the batch and their device stage tables
        host["aug_tea"], host["aug_stu"] = S.aug_params(b, seed + 6), S.aug_params(b, seed + 7, shear_y=True)
        inp.theta_tea = RW.stage_table(RW.recon_stages(host["aug_tea"], 4.0, b), 64, 64, torch.float32, None)[0].to(dev)
        inp.theta_stu = RW.stage_table(RW.recon_stages(host["aug_stu"], 4.0, b), 64, 64, torch.float16, torch.float16)[0].to(dev)
    return host, inp


def _oracle_step(host, a1, a2, teacher, student, scale=65536.0):
    t1 = R.adain_mix(host["feat_src"], host["feat_tgt_ori"], a1)
    t2 = R.adain_mix(host["feat_tgt_tea"], host["feat_src_ori"], a2)
    y_t_tea = host["y_t_tea"]
    if "aug_tea" in host:
        y_t_tea = R.teacher_recon([y_t_tea], [host["aug_tea"]], 4.0)              # train_human.py:359-372
    conf, pos, table = R.confidence_mask(y_t_tea, 0.9)
    mask, thresh, act = R.consistency_mask(y_t_tea, 0.5)
    rect = R.rectify(y_t_tea, 2)
    y_s = host["y_s"].float().clone().requires_grad_(True)
    if "aug_stu" in host:
        y_t = host["y_t_stu"].clone().requires_grad_(True)                        # fp16 leaf, autocast re-warp
        y_t_recon = R.student_recon(y_t, host["aug_stu"], 4.0)                    # :417-423
    else:
        y_t = host["y_t_stu"].float().clone().requires_grad_(True)
        y_t_recon = y_t
    loss_s = R.joints_mse_loss(y_s, host["label_s"], host["weight_s"])
    loss_c = R.cons_loss(y_t_recon.float(), rect, tea_mask=mask)
    ((loss_s + 1.0 * loss_c) * scale).backward()
    R.ema_step(teacher, student, 0.999)
    hits, valid, pred = R.pck_counts(host["y_s"].numpy(), host["label_s"].numpy())
    return dict(y_t_tea=y_t_tea, y_t_stu_recon=y_t_recon.detach(), t_s2t=t1, t_t2s=t2, conf_table=table, position=pos, tea_mask=mask, rectified=rect,
                loss_s=loss_s.detach(), loss_c=loss_c.detach(), grad_y_s=y_s.grad, grad_y_t_stu=y_t.grad.float(),
                hits=hits, valid=valid, pred=pred)


def _check(out, ref):
    if out["y_t_tea_recon"] is not None:      # (None with fuse_teacher_decode: the map is decoded where it is gathered)
        assert torch.equal(out["y_t_tea_recon"].cpu(), ref["y_t_tea"])          # gathers: bit-exact
    assert torch.equal(out["y_t_stu_recon"].cpu(), ref["y_t_stu_recon"].to(out["y_t_stu_recon"].dtype))
    assert_close_scaled(out["t_s2t"], ref["t_s2t"], 1e-5, "s2t")
    assert_close_scaled(out["t_t2s"], ref["t_t2s"], 1e-5, "t2s")
    assert torch.equal(out["conf_table"].cpu(), ref["conf_table"])
    assert torch.equal(out["position"].cpu(), ref["position"])
    assert torch.equal(out["tea_mask"].cpu(), ref["tea_mask"])
    if out["rectified"] is not None:  # unfused route materialises rectify(y_t_tea)
        assert_close_scaled(out["rectified"], ref["rectified"], 1e-5, "rectified")
    np.testing.assert_array_equal(out["tea_preds"].cpu().numpy(), R.get_max_preds_torch(ref["y_t_tea"])[0].numpy())
    assert_close_scaled(out["loss_s"], ref["loss_s"], 1e-5, "loss_s")
    assert_close_scaled(out["loss_c"], ref["loss_c"], 1e-5, "loss_c")
    assert_close_scaled(out["loss_all"], ref["loss_s"] + ref["loss_c"], 1e-5, "loss_all")
    assert_close_scaled(out["grad_y_s"].float(), ref["grad_y_s"], 1e-2, "grad y_s (fp16)")
    assert_close_scaled(out["grad_y_t_stu"].float(), ref["grad_y_t_stu"], 1e-2, "grad y_t_stu (fp16)")
    np.testing.assert_array_equal(out["pck_counts"].cpu().numpy(), np.stack([ref["hits"], ref["valid"]]))
    np.testing.assert_array_equal(out["pred"].cpu().numpy(), ref["pred"])


@pytest.mark.parametrize("rewarp", [True, False])
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("parallel", [False, True])
def test_step_eager_vs_oracle(dev, parallel, fused, rewarp):
    host, inp = _inputs(dev, rewarp=rewarp)
    shapes = [(64, 3, 7, 7), (64,), (17,), (256, 64, 1, 1), (5000,)]
    s_cpu, t_cpu = S.parameter_list(shapes, 1), S.parameter_list(shapes, 2)
    student, teacher = Bag(s_cpu).to(dev), Bag(t_cpu).to(dev)
    step = HotPathStep(teacher, student, sigma=2, parallel=parallel, fused=fused)
    R.ema_init(t_cpu, s_cpu)
    for it in range(2):
        out = step.run(inp)
        torch.cuda.synchronize()
        ref = _oracle_step(host, 0.3, 0.8, t_cpu, s_cpu)
        _check(out, ref)
        for p, e in zip(teacher.parameters(), t_cpu):
            assert torch.equal(p.detach().cpu(), e)


@pytest.mark.parametrize("fused,ema_parallel", [(True, True), (True, False), (False, True)])
def test_step_graph_replay_with_device_alpha(dev, fused, ema_parallel):
    host, inp = _inputs(dev, seed=11)
    shapes = [(128, 64, 3, 3), (128,), (33,)]
    s_cpu, t_cpu = S.parameter_list(shapes, 3), S.parameter_list(shapes, 4)
    student, teacher = Bag(s_cpu).to(dev), Bag(t_cpu).to(dev)
    step = HotPathStep(teacher, student, sigma=2, fused=fused, ema_parallel=ema_parallel)
    R.ema_init(t_cpu, s_cpu)
    step.capture(inp, include_ema=True, warmup=2)  # 2 eager warm-up steps + the capture pass do not replay
    for _ in range(2):
        R.ema_step(t_cpu, s_cpu, 0.999)
    for a1, a2 in ((0.3, 0.8), (1.0, 0.0), (0.55, 0.45)):
        inp.alpha_s2t.fill_(a1)
        inp.alpha_t2s.fill_(a2)
        # new inputs in the same static buffers
        inp.y_s.copy_(torch.roll(inp.y_s, 1, dims=0))
        host["y_s"] = torch.roll(host["y_s"], 1, dims=0)
        # new augmentation parameters: the stage tables are device tensors the graph reads in place
        host["aug_tea"], host["aug_stu"] = S.aug_params(4, int(a1 * 100)), S.aug_params(4, int(a2 * 100) + 1)
        inp.theta_tea.copy_(RW.stage_table(RW.recon_stages(host["aug_tea"], 4.0, 4), 64, 64, torch.float32, None)[0])
        inp.theta_stu.copy_(RW.stage_table(RW.recon_stages(host["aug_stu"], 4.0, 4), 64, 64, torch.float16, torch.float16)[0])
        out = step.replay()
        torch.cuda.synchronize()
        ref = _oracle_step(host, a1, a2, t_cpu, s_cpu)
        _check(out, ref)
        for p, e in zip(teacher.parameters(), t_cpu):
            assert torch.equal(p.detach().cpu(), e)


def test_step_graph_alpha_feed(dev):
    """alpha_feed runs inside the captured graph behind the AdaIN launches and loads the NEXT replay's scalars
    from a device table (what bench.py uses): replay r must see row r of the table."""
    host, inp = _inputs(dev, seed=21, rewarp=False)
    shapes = [(64,)]
    student, teacher = Bag(S.parameter_list(shapes, 3)).to(dev), Bag(S.parameter_list(shapes, 4)).to(dev)
    step = HotPathStep(teacher, student, sigma=2)
    table = torch.tensor([[0.1, 0.9], [0.5, 0.25], [1.0, 0.0], [0.0, 1.0], [0.7, 0.3], [0.2, 0.6]], device=dev)
    pair = torch.zeros(2, device=dev)
    inp.alpha_s2t, inp.alpha_t2s = pair[0:1], pair[1:2]
    row = torch.ones(1, dtype=torch.int64, device=dev)

    def feed():
        pair.copy_(table.index_select(0, row).view(2))
        row.add_(1).remainder_(table.shape[0])

    step.alpha_feed = feed
    pair.copy_(table[0])
    step.capture(inp, include_ema=False, warmup=2)     # the two eager warm-up passes consume rows 0 and 1
    for r in (2, 3, 4):
        out = step.replay()
        torch.cuda.synchronize()
        a1, a2 = float(table[r, 0]), float(table[r, 1])
        assert_close_scaled(out["t_s2t"], R.adain_mix(host["feat_src"], host["feat_tgt_ori"], a1), 1e-5, f"s2t replay row {r}")
        assert_close_scaled(out["t_t2s"], R.adain_mix(host["feat_tgt_tea"], host["feat_src_ori"], a2), 1e-5, f"t2s replay row {r}")


def test_step_bytes_accounting(dev):
    _, inp = _inputs(dev, b=2, k=16)
    for fused in (True, False):
        by = step_algorithmic_bytes(inp, n_params=1000, fused=fused)
        assert by["total"] == sum(v for k_, v in by.items() if k_ != "total")
    assert by["ema"] == 12000 and by["adain_mix"] == 2 * 3 * inp.feat_src.numel() * 4
    assert by["total"] == sum(v for k_, v in by.items() if k_ != "total")


def _c2_inputs(dev, seed=41, k_views=2):
    """BASELINE.json configs[1] shapes (batch 32, 16 keypoints; 512x32x32 features are cut to 64 channels to keep
    the CPU oracle quick — planes are independent, the full tensor is covered by test_gpu_oracle), k teacher views
    each with its own augmentation, and the student's target images for the occlusion stage."""
    b, k = S.CONFIGS["C2"]["batch"], S.CONFIGS["C2"]["joints"]
    host, inp = _inputs(dev, b=b, k=k, n_feat_c=64, seed=seed, rewarp=True)
    # peaks up to 2.4: the mean over two views that disagree on the position still clears occlude_thresh = 0.9 somewhere
    views = [S.heatmaps(b, k, seed + 20 + v, peak=(1.0, 2.4)) for v in range(k_views)]
    augs = [S.aug_params(b, seed + 30 + v) for v in range(k_views)]
    host["y_t_teas"], host["aug_teas"] = views, augs
    inp.y_t_tea = [v.to(dev) for v in views]
    inp.theta_tea = [RW.stage_table(RW.recon_stages(a, 4.0, b), 64, 64, torch.float32, None)[0].to(dev) for a in augs]
    g = torch.Generator().manual_seed(seed + 40)
    host["x_t_stu"] = torch.randn(b, 3, 256, 256, generator=g)
    return host, inp


def _oracle_c2(host, a1, a2, teacher, student, occlusion_seed=None):
    h = dict(host)
    h["y_t_tea"] = R.teacher_recon(host["y_t_teas"], host["aug_teas"], 4.0)      # train_human.py:359-372, k views
    h.pop("aug_tea", None)
    ref = _oracle_step(h, a1, a2, teacher, student)
    if occlusion_seed is not None:
        ref["x_t_stu"] = R.occlude_keypoints(host["x_t_stu"], ref["conf_table"], ref["position"].numpy(), host["aug_stu"], 4.0,
                                             0.5, 10, 256, rng=np.random.RandomState(occlusion_seed))    # :385-412
    return ref


def test_step_at_c2_size_with_k2_views_and_occlusion_eager(dev):
    """One eager step at the BASELINE config's batch / keypoint count with k = 2 teacher views (mean over the
    re-warped views) and the occlusion paste driven by the reference's np.random draws."""
    host, inp = _c2_inputs(dev)
    inp.x_t_stu, inp.aug_param_stu = host["x_t_stu"].to(dev), host["aug_stu"]
    shapes = [(64, 3, 7, 7), (64,), (17,), (256, 64, 1, 1), (5000,)]
    s_cpu, t_cpu = S.parameter_list(shapes, 1), S.parameter_list(shapes, 2)
    student, teacher = Bag(s_cpu).to(dev), Bag(t_cpu).to(dev)
    step = HotPathStep(teacher, student, sigma=2, rng=np.random.RandomState(77))
    R.ema_init(t_cpu, s_cpu)
    out = step.run(inp)
    torch.cuda.synchronize()
    ref = _oracle_c2(host, 0.3, 0.8, t_cpu, s_cpu, occlusion_seed=77)
    _check(out, ref)
    assert out["x_t_stu"] is not None and not torch.equal(out["x_t_stu"].cpu(), host["x_t_stu"]), "no sample was occluded"
    assert torch.equal(out["x_t_stu"].cpu(), ref["x_t_stu"])                     # gather + paste: bit-exact
    for p, e in zip(teacher.parameters(), t_cpu):
        assert torch.equal(p.detach().cpu(), e)


def test_step_at_c2_size_with_k2_views_graph(dev):
    """The same step (without the host-driven occlusion stage) captured once and replayed with new inputs."""
    host, inp = _c2_inputs(dev, seed=43)
    shapes = [(128, 64, 3, 3), (128,), (33,)]
    s_cpu, t_cpu = S.parameter_list(shapes, 3), S.parameter_list(shapes, 4)
    student, teacher = Bag(s_cpu).to(dev), Bag(t_cpu).to(dev)
    step = HotPathStep(teacher, student, sigma=2)
    R.ema_init(t_cpu, s_cpu)
    step.capture(inp, include_ema=True, warmup=2)
    for _ in range(2):
        R.ema_step(t_cpu, s_cpu, 0.999)
    b = S.CONFIGS["C2"]["batch"]
    for rep in range(2):
        host["y_t_teas"] = [torch.roll(v, rep + 1, dims=0) for v in host["y_t_teas"]]
        for dst, src in zip(inp.y_t_tea, host["y_t_teas"]):
            dst.copy_(src)
        host["aug_teas"] = [S.aug_params(b, 500 + 10 * rep + v) for v in range(2)]
        for dst, a in zip(inp.theta_tea, host["aug_teas"]):
            dst.copy_(RW.stage_table(RW.recon_stages(a, 4.0, b), 64, 64, torch.float32, None)[0])
        out = step.replay()
        torch.cuda.synchronize()
        ref = _oracle_c2(host, 0.3, 0.8, t_cpu, s_cpu)
        _check(out, ref)
        for p, e in zip(teacher.parameters(), t_cpu):
            assert torch.equal(p.detach().cpu(), e)


@pytest.mark.parametrize("graph", [False, True])
def test_step_with_the_teacher_chain_in_one_launch(dev, graph):
    """fuse_teacher_decode: teacher re-warp + decode + conf_table + k-th value mask as ONE launch (the re-warped map is
    never written) at the BASELINE config's size with one teacher view — against the oracle, and bit for bit against the
    step that materialises the map; eagerly (with the occlusion stage, which only needs conf_table / position) and as a
    replayed graph."""
    host, inp = _c2_inputs(dev, seed=47, k_views=1)
    if not graph:
        inp.x_t_stu, inp.aug_param_stu = host["x_t_stu"].to(dev), host["aug_stu"]
    shapes = [(64, 3, 7, 7), (64,), (17,)]
    outs = {}
    for fuse in (False, True):
        s_cpu, t_cpu = S.parameter_list(shapes, 1), S.parameter_list(shapes, 2)
        student, teacher = Bag(s_cpu).to(dev), Bag(t_cpu).to(dev)
        step = HotPathStep(teacher, student, sigma=2, rng=np.random.RandomState(78), fuse_teacher_decode=fuse)
        R.ema_init(t_cpu, s_cpu)
        if graph:
            step.capture(inp, include_ema=True, warmup=1)
            R.ema_step(t_cpu, s_cpu, 0.999)
            out = step.replay()
        else:
            out = step.run(inp)
        torch.cuda.synchronize()
        assert (out["y_t_tea_recon"] is None) == fuse
        outs[(fuse, 'kernels')] = step.kernels_per_step
        outs[fuse] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in out.items()}
        ref = _oracle_c2(host, 0.3, 0.8, t_cpu, s_cpu, occlusion_seed=None if graph else 78)
        _check(out, ref)
        if not graph:
            assert torch.equal(out["x_t_stu"].cpu(), ref["x_t_stu"])
    for k in ("conf_table", "position", "tea_mask", "mask_thresh", "tea_preds", "loss_all", "loss_s", "loss_c", "grad_y_s",
              "grad_y_t_stu", "pck_counts"):
        assert torch.equal(outs[True][k], outs[False][k]), k
    assert outs[(False, 'kernels')] - outs[(True, 'kernels')] == 1
    by_two = step_algorithmic_bytes(inp, n_params=1000, fused=True)
    by_one = step_algorithmic_bytes(inp, n_params=1000, fused=True, fuse_teacher_decode=True)
    hm = inp.teacher_views[0].numel() * 4
    assert by_two["total"] - by_one["total"] == 2 * hm      # the map's write and its re-read


def test_ticket_words_are_zero_between_launches(dev):
    """Self-resetting last-CTA tickets: after any number of launches every word of the pool is zero again; a word
    left dirty (a launch that aborted mid-grid) is reported by check_tickets() and cleared by reset=True."""
    from uda_poseestimation_b200 import _lib
    host, inp = _inputs(dev, seed=31)
    student, teacher = Bag(S.parameter_list([(64,)], 3)).to(dev), Bag(S.parameter_list([(64,)], 4)).to(dev)
    step = HotPathStep(teacher, student, sigma=2)
    for _ in range(3):
        step.run(inp)
    U.accuracy(inp.y_s, inp.label_s)
    U.check_tickets()                                   # all zero
    pool = _lib._ticket_pools[dev.index if dev.index is not None else torch.cuda.current_device()]
    pool.buf[pool.SLOTS - 1] = 7                        # what an aborted launch would leave behind
    with pytest.raises(U.UdapeError, match="ticket"):
        U.check_tickets()
    U.check_tickets(reset=True)
    U.check_tickets()
    out = step.run(inp)                                 # and the pool keeps working
    torch.cuda.synchronize()
    assert torch.isfinite(out["loss_all"]).all()

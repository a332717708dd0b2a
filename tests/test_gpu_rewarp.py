"""GPU parity of the batched affine re-warp (csrc/rewarp.cu) — the trainers' per-sample tF.affine
loops (train_human.py:359-372, :385-412, :417-423) — against

* the golden fixtures produced by those loops running through torchvision on the CPU
  (tests/golden/make_golden.py::golden_rewarp), and
* the oracle's loops (oracle/reference_port.py, torchvision on the CPU) at BASELINE config sizes.

Bars: forward values are gathered, never computed — BIT-EXACT (every source index must match the
reference's nearest-neighbour choice, including ties and the half-precision grids of the autocast
block); gradients are float32 sums rounded once — 1e-5 relative in fp32, 1e-2 in fp16 (2e-2 in
bf16, see tests/test_oracle_golden.py::test_student_recon).
"""
import numpy as np
import pytest
import torch

import uda_poseestimation_b200 as U
from conftest import assert_close_scaled
from oracle import reference_port as R
from uda_poseestimation_b200 import rewarp as RW
from uda_poseestimation_b200 import synthetic as S

pytestmark = pytest.mark.gpu
GRAD_TOL = {torch.float32: 1e-5, torch.float16: 1e-2, torch.bfloat16: 2e-2}


def C(a, dev):
    return torch.from_numpy(np.asarray(a)).to(dev)


def aug_of(a):
    a = torch.from_numpy(np.asarray(a))
    return [a[:, 0], [a[:, 1].long(), a[:, 2].long()], [a[:, 3], a[:, 4]], a[:, 5]]


@pytest.mark.parametrize("tag,k", [("tea1", 1), ("tea2", 2), ("teaodd", 1)])
def test_teacher_recon_golden(golden, dev, tag, k):
    g = golden("rewarp")
    views = [C(g[f"{tag}_in{i}"], dev) for i in range(k)]
    augs = [aug_of(g[f"{tag}_v{i}_aug"]) for i in range(k)]
    out = U.teacher_recon(views, augs, 4.0)
    assert out.dtype == torch.float32 and out.shape == views[0].shape
    assert torch.equal(out.cpu(), torch.from_numpy(g[f"{tag}_out"]))


@pytest.mark.parametrize("tag,dt", [("stu16", torch.float16), ("stubf", torch.bfloat16), ("stu32", torch.float32)])
def test_student_recon_golden(golden, dev, tag, dt):
    g = golden("rewarp")
    y = C(g[f"{tag}_in"], dev).to(dt).requires_grad_(True)
    aug = aug_of(g[f"{tag}_aug"])
    ac = None if dt == torch.float32 else dt
    out = U.student_recon(y, aug, 4.0, autocast=ac)
    assert out.dtype == dt
    assert torch.equal(out.detach().float().cpu(), torch.from_numpy(g[f"{tag}_out"]))
    (out.float() * C(g[f"{tag}_G"], dev)).sum().backward()
    assert y.grad.dtype == dt
    assert_close_scaled(y.grad.float(), g[f"{tag}_grad"], GRAD_TOL[dt], f"{tag} grad")
    if ac is not None:
        # inside the trainers' autocast block the dtype is picked up from the context
        with torch.autocast("cuda", dtype=dt):
            out2 = U.student_recon(y.detach(), aug, 4.0)
        assert torch.equal(out2, out.detach())
        with pytest.raises(NotImplementedError):
            U.student_recon(y.detach(), aug, 4.0)  # a half tensor outside autocast


def test_occlusion_golden(golden, dev):
    g = golden("rewarp")
    ratio, rate, size, image = g["occ_args"]
    rng = np.random.RandomState(int(g["occ_seed"]))
    out = U.occlude_keypoints(C(g["occ_in"], dev), g["occ_conf_table"], g["occ_pred_position"], aug_of(g["occ_aug"]),
                              float(ratio), float(rate), int(size), int(image), rng=rng)
    assert torch.equal(out.cpu(), torch.from_numpy(g["occ_out"]))
    # no sample selected: the batch comes back unchanged (a copy)
    x = C(g["occ_in"], dev)
    same = U.occlude_keypoints(x, g["occ_conf_table"], g["occ_pred_position"], aug_of(g["occ_aug"]), float(ratio), -1.0,
                               int(size), int(image), rng=np.random.RandomState(0))
    assert torch.equal(same, x) and same.data_ptr() != x.data_ptr()


@pytest.mark.parametrize("cfg", ["C2", "C4"])
def test_recon_vs_oracle_at_config_sizes(dev, cfg):
    b, k = S.CONFIGS[cfg]["batch"], S.CONFIGS[cfg]["joints"]
    tea = S.heatmaps(b, k, seed=31, peak=(0.3, 1.2))
    aug_t = S.aug_params(b, seed=32)
    assert torch.equal(U.teacher_recon([tea.to(dev)], [aug_t], 4.0).cpu(), R.teacher_recon([tea], [aug_t], 4.0))
    stu = S.heatmaps(b, k, seed=33).half()
    aug_s = S.aug_params(b, seed=34, shear_y=True)
    y = stu.to(dev).requires_grad_(True)
    out = U.student_recon(y, aug_s, 4.0, autocast=torch.float16)
    y_ref = stu.clone().requires_grad_(True)
    ref = R.student_recon(y_ref, aug_s, 4.0)
    assert torch.equal(out.detach().cpu(), ref.detach())
    G = torch.randn(b, k, 64, 64, generator=torch.Generator().manual_seed(35)).half()
    out.backward(G.to(dev))
    ref.backward(G)
    assert_close_scaled(y.grad.float(), y_ref.grad.float(), 1e-2, "student recon grad")


@pytest.mark.parametrize("shape", [(2, 3, 37, 53), (1, 1, 1, 1), (3, 2, 16, 128), (2, 3, 256, 256)])
def test_affine_nearest_vs_torchvision(dev, shape):
    """One-stage batched tF.affine, ragged sizes (scalar path), image-sized planes."""
    from torchvision.transforms import functional as tF

    rng = np.random.RandomState(shape[2])
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1))
    b = shape[0]
    ang = [float(rng.uniform(-180, 180)) for _ in range(b)]
    sc = [float(rng.uniform(0.5, 1.7)) for _ in range(b)]
    sh = [[float(rng.uniform(-30, 30)), float(rng.uniform(-30, 30))] for _ in range(b)]
    tr = [[float(rng.randint(-13, 14)) / 4, float(rng.randint(-13, 14)) / 4] for _ in range(b)]
    out = U.affine_nearest(x.to(dev), ang, tr, sc, sh)
    for i in range(b):
        assert torch.equal(out[i].cpu(), tF.affine(x[i], ang[i], translate=tr[i], scale=sc[i], shear=sh[i])), i
    one = U.affine_nearest(x[0].to(dev), ang[0], tr[0], sc[0], sh[0])
    assert one.shape == x[0].shape and torch.equal(one, out[0])


def test_rewarp_properties_full_size(dev):
    """Size-independent properties at the C5-class size (256x21x64x64)."""
    b, k = 256, 21
    y = torch.randn(b, k, 64, 64, device=dev)
    ident = [torch.zeros(b, dtype=torch.float64), [torch.zeros(b, dtype=torch.int64)] * 2,
             [torch.zeros(b, dtype=torch.float64)] * 2, torch.ones(b, dtype=torch.float64)]
    assert torch.equal(U.teacher_recon([y], [ident], 4.0), y)                      # identity parameters
    shift = [torch.zeros(b, dtype=torch.float64), [torch.full((b,), 8, dtype=torch.int64), torch.full((b,), -12, dtype=torch.int64)],
             [torch.zeros(b, dtype=torch.float64)] * 2, torch.ones(b, dtype=torch.float64)]
    out = U.teacher_recon([y], [shift], 4.0)                                       # +2 px in x, -3 px in y
    want = torch.zeros_like(y)
    want[:, :, :61, 2:] = y[:, :, 3:, :62]
    assert torch.equal(out, want)
    aug = S.aug_params(b, seed=7, shear_y=True)
    table, half_mask, _ = RW.stage_table(RW.recon_stages(aug, 4.0, b), 64, 64, torch.float32, None)
    theta = table.to(dev)
    a_, b_ = torch.randn_like(y), torch.randn_like(y)
    ra, rb = RW.gather(a_, theta), RW.gather(b_, theta)
    assert torch.equal(RW.gather(a_ + b_, theta), ra + rb)                         # a gather is linear, exactly
    # adjoint: <rewarp(x), g> == <x, rewarp^T(g)>, and the backward is deterministic
    x = a_.clone().requires_grad_(True)
    g = b_
    RW.gather(x, theta).backward(g)
    lhs = (ra.double() * g.double()).sum()
    rhs = (a_.double() * x.grad.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-7 * float(ra.double().norm() * g.double().norm())
    x2 = a_.clone().requires_grad_(True)
    RW.gather(x2, theta).backward(g)
    assert torch.equal(x.grad, x2.grad)


def test_rewarp_graph_capture(dev):
    """theta is a device tensor: a captured re-warp follows fresh augmentation parameters."""
    b, k = 8, 4
    y = torch.randn(b, k, 64, 64, device=dev)
    tables = [RW.stage_table(RW.recon_stages(S.aug_params(b, seed=s), 4.0, b), 64, 64, torch.float32, None)[0] for s in (1, 2)]
    theta = tables[0].to(dev)
    eager = [RW.gather(y, t.to(dev)) for t in tables]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        RW.gather(y, theta)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = RW.gather(y, theta)
    for t, want in zip(tables, eager):
        theta.copy_(t)
        graph.replay()
        assert torch.equal(out, want)


def test_rewarp_argument_errors(dev):
    y = torch.randn(2, 3, 8, 8, device=dev)
    theta = torch.zeros(2, 3, 6, device=dev)
    with pytest.raises(RuntimeError):
        RW.gather(y.cpu(), theta)                       # no CPU fallback
    with pytest.raises(ValueError):
        RW.gather(y, torch.zeros(3, 3, 6, device=dev))  # batch mismatch
    with pytest.raises(ValueError):
        RW.gather(y, torch.zeros(2, 5, 6, device=dev))  # too many stages
    with pytest.raises(ValueError):
        U.teacher_recon([y, y], [None], 4.0)
    big = torch.randn(1, 1, 256, 256, device=dev, requires_grad=True)
    with pytest.raises(ValueError):                      # backward keeps the inverted map in shared memory
        RW.gather(big, torch.zeros(1, 1, 6, device=dev)).sum().backward()


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
def test_rewarp_routes_agree(dev, dt, monkeypatch):
    """The shared-memory staged route and the global-gather route are the same function: identical
    bits forward (1 and 3 views) and backward, including a plane smaller than one pass of the CTA."""
    for (b, k, h, w) in [(5, 7, 64, 64), (3, 2, 32, 32), (2, 3, 24, 40)]:
        ys = [(torch.randn(b, k, h, w, device=dev) * 3).to(dt) for _ in range(3)]
        augs = [S.aug_params(b, seed=50 + i, shear_y=True) for i in range(3)]
        ac = None if dt == torch.float32 else dt
        tabs = [RW.stage_table(RW.recon_stages(a_, 4.0, b), h, w, dt, ac) for a_ in augs]
        thetas = [t[0].to(dev) for t in tabs]
        half_mask, code = tabs[0][1], tabs[0][2]
        g = (torch.randn(b, k, h, w, device=dev)).to(dt)
        res = {}
        for route in ("smem", "global"):
            if route == "global":
                monkeypatch.setenv("UDAPE_REWARP_GLOBAL", "1")
            else:
                monkeypatch.delenv("UDAPE_REWARP_GLOBAL", raising=False)
            one = RW._launch_fwd([ys[0]], [thetas[0]], half_mask, code, torch.empty_like(ys[0]))
            three = RW._launch_fwd(ys, thetas, half_mask, code, torch.empty_like(ys[0]))
            x = ys[0].clone().requires_grad_(True)
            RW._Rewarp.apply(x, thetas[0], half_mask, code).backward(g)   # inverse plan when the plane qualifies
            noplan = RW._launch_bwd(g, thetas[0], half_mask, code, None)   # the backward inverts the map itself
            res[route] = (one, three, x.grad, noplan)
            assert torch.equal(x.grad, noplan)
        monkeypatch.delenv("UDAPE_REWARP_GLOBAL", raising=False)
        for r_s, r_g in zip(res["smem"], res["global"]):
            assert torch.equal(r_s, r_g), (b, k, h, w)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
def test_rewarp_ring_depths_agree(dev, dt, monkeypatch):
    """A CTA that owns six or more planes (one CTA per sample at batch 32) stages them through a six-buffer
    ring instead of three: same bits, also for channel counts that are not a multiple of the ring."""
    monkeypatch.setenv("UDAPE_REWARP_WIDE", "0")      # the ring depths belong to the 256-thread kernels
    for (b, k) in [(32, 16), (33, 7), (40, 21), (148, 6)]:
        y = (torch.randn(b, k, 64, 64, device=dev) * 3).to(dt)
        ac = None if dt == torch.float32 else dt
        tab = RW.stage_table(RW.recon_stages(S.aug_params(b, seed=90 + k, shear_y=True), 4.0, b), 64, 64, dt, ac)
        theta = tab[0].to(dev)
        monkeypatch.setenv("UDAPE_REWARP_RING", "3")
        shallow = RW._launch_fwd([y], [theta], tab[1], tab[2], torch.empty_like(y))
        monkeypatch.delenv("UDAPE_REWARP_RING")
        deep = RW._launch_fwd([y], [theta], tab[1], tab[2], torch.empty_like(y))
        assert torch.equal(shallow, deep), (b, k)


def test_rewarp_backward_long_lists(dev):
    """Zoom factors above ~1.7 give source pixels with more than four contributors (the backward kernel
    keeps four in registers and loops over the rest): checked against autograd through torchvision."""
    b, k = 4, 3
    aug = [torch.tensor([10.0, -35.0, 80.0, 0.0], dtype=torch.float64),
           [torch.tensor([3, -5, 0, 8]), torch.tensor([-2, 7, 0, 1])],
           [torch.tensor([5.0, 0.0, -12.0, 0.0], dtype=torch.float64), torch.zeros(4, dtype=torch.float64)],
           torch.tensor([2.9, 2.2, 3.5, 1.0], dtype=torch.float64)]
    for dt, tol in ((torch.float32, 1e-5), (torch.float16, 1e-2)):
        x = S.heatmaps(b, k, seed=61).to(dt)
        G = torch.randn(b, k, 64, 64, generator=torch.Generator().manual_seed(62)).to(dt)
        ac = None if dt == torch.float32 else dt
        y = x.to(dev).requires_grad_(True)
        out = U.student_recon(y, aug, 4.0, autocast=ac)
        out.backward(G.to(dev))
        y_ref = x.clone().requires_grad_(True)
        ref = R.student_recon(y_ref, aug, 4.0)
        ref.backward(G)
        assert torch.equal(out.detach().cpu(), ref.detach())
        assert_close_scaled(y.grad.float(), y_ref.grad.float(), tol, f"long-list grad {dt}")
        # more than four output pixels really do share a source pixel here
        src = R.recon_source_index(10.0, 3, -2, 5.0, 0.0, 2.9, 4.0, 64, 64, dt, ac)
        assert np.bincount(src[src >= 0]).max() > 4


@pytest.mark.parametrize("cluster", ["1", "2", "4", "8"])
def test_rewarp_cluster_kernels_under_contention(dev, cluster, monkeypatch):
    """Every cluster size gives the same bits (the default policy picks one CTA per sample at this batch
    size; smaller batches split a sample's channels over a cluster of 2, 4 or 8 CTAs).
    The CTAs of a cluster exchange the composed map / the inverted lists through distributed shared
    memory; when other kernels compete for the SMs the CTAs of a cluster drift apart in time, which is
    what exposes a missing cluster barrier.  Forward and backward are repeated on two high-priority
    streams next to a bandwidth hog and must reproduce the quiet result bit for bit."""
    b, k = 32, 16
    y = S.heatmaps(b, k, seed=71).to(dev).half()
    g = torch.randn(b, k, 64, 64, device=dev).half()
    t = RW.stage_table(RW.recon_stages(S.aug_params(b, seed=72, shear_y=True), 4.0, b), 64, 64, torch.float16, torch.float16)
    theta = t[0].to(dev)
    monkeypatch.setenv("UDAPE_REWARP_GLOBAL", "1")      # reference bits: the route without clusters
    ref_f = RW.gather(y, theta, t[1], torch.float16)
    ref_b = RW.gather_backward(g, theta, t[1], torch.float16)
    monkeypatch.delenv("UDAPE_REWARP_GLOBAL")
    monkeypatch.setenv("UDAPE_REWARP_WIDE", "0")        # the cluster kernels are the 256-thread route
    monkeypatch.setenv("UDAPE_REWARP_CLUSTER", cluster)
    plans = [RW.inverse_plan_buffer(y) for _ in range(2)]
    torch.cuda.synchronize()
    hog_a, hog_b = torch.empty(64 << 20, device=dev), torch.empty(64 << 20, device=dev)
    streams = [torch.cuda.Stream(dev, priority=-1) for _ in range(2)]
    hog = torch.cuda.Stream(dev)
    outs = []
    for it in range(40):
        with torch.cuda.stream(hog):
            hog_b.copy_(hog_a)
        for s_, plan in zip(streams, plans):
            with torch.cuda.stream(s_):
                outs.append((RW.gather(y, theta, t[1], torch.float16), RW.gather_backward(g, theta, t[1], torch.float16)))
                # the plan route: forward + cluster-built inverse plan, then the plan-based backward
                outs.append((RW.gather(y, theta, t[1], torch.float16, plan=plan),
                             RW.gather_backward(g, theta, t[1], torch.float16, plan=plan)))
    torch.cuda.synchronize()
    for f, bw in outs:
        assert torch.equal(f, ref_f) and torch.equal(bw, ref_b)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
def test_rewarp_wide_route_equals_the_256_thread_kernels(dev, dt, monkeypatch):
    """The wide forward route (one 512-thread CTA per sample, compact map builder) against the 256-thread cluster
    kernels and the global-memory route, and the backward with / without an inverse plan on every route: identical
    bits for full and small planes, odd batch / channel counts, and zoom factors that give long contributor lists."""
    for (b, k, h, w) in [(37, 21, 64, 64), (5, 3, 32, 32), (3, 2, 24, 40), (2, 5, 16, 8)]:
        x = (torch.randn(b, k, h, w, generator=torch.Generator().manual_seed(b)) * 3).to(dt).to(dev)
        g = torch.randn(b, k, h, w, generator=torch.Generator().manual_seed(b + 1)).to(dt).to(dev)
        aug = S.aug_params(b, seed=130 + b, shear_y=True, scale=(0.3, 1.6))      # 1 / 0.3: lists of ten and more
        half = dt != torch.float32
        table, mask, _ = RW.stage_table(RW.recon_stages(aug, 4.0, b), h, w, dt, dt if half else None)
        theta = table.to(dev)
        gd = dt if half else None
        res = {}
        for route, env in (("wide", {}), ("cta256", {"UDAPE_REWARP_WIDE": "0"}), ("global", {"UDAPE_REWARP_GLOBAL": "1"})):
            for name in ("UDAPE_REWARP_WIDE", "UDAPE_REWARP_GLOBAL"):
                monkeypatch.delenv(name, raising=False)
            for name, val in env.items():
                monkeypatch.setenv(name, val)
            res[route] = (RW.gather(x, theta, mask, gd), RW.gather_backward(g, theta, mask, gd))
        monkeypatch.setenv("UDAPE_REWARP_WIDE", "0")
        plan = RW.build_inverse_plan(x, theta, mask, gd)
        if plan is not None:
            res["plan"] = (res["wide"][0], RW.gather_backward(g, theta, mask, gd, plan=plan))
        monkeypatch.delenv("UDAPE_REWARP_WIDE")
        for route, (f, bw) in res.items():
            assert torch.equal(f, res["wide"][0]), (route, "forward", b, k, h, w)
            assert torch.equal(bw, res["wide"][1]), (route, "backward", b, k, h, w)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("c", [5, 8])
def test_rewarp_backward_degenerate_maps(dev, dt, c):
    """The push plan on maps it was not tuned for, against a sequential float32 scatter on the host (ascending
    output pixel, one rounding): every output pixel onto ONE source pixel (a list of 4096 entries: 4095 tail
    slots, one group), no output pixel inside the image (nothing is pushed), the identity (no group), a 2x zoom
    (every source pixel four contributors) and an ordinary augmentation; an odd channel count leaves the last
    pass of a sample partly empty.  Also: the plan route and the list route agree bit for bit."""
    h = w = 64
    hw = h * w
    one = [[0.0, 0.0, 0.1], [0.0, 0.0, -0.2]]          # grid constant: every pixel reads the same source
    none = [[0.0, 0.0, 5.0], [0.0, 0.0, 5.0]]          # grid outside [-1, 1]
    ident = [[1.0 / 32, 0.0, 0.0], [0.0, 1.0 / 32, 0.0]]   # rows are theta / (0.5 * size), pixels in, grid out
    zoom = [[0.5 / 32, 0.0, 0.0], [0.0, 0.5 / 32, 0.0]]
    theta = torch.tensor([one, none, ident, zoom], dtype=torch.float32).reshape(4, 1, 6)
    aug = S.aug_params(1, seed=404)
    half = dt != torch.float32
    extra = RW.stage_table(RW.recon_stages(aug, 4.0, 1), h, w, torch.float32, None)[0][:, :1]   # one ordinary stage
    theta = torch.cat([theta, extra.reshape(1, 1, 6)], 0).to(dev)
    b = theta.shape[0]
    # the source index of every output pixel, from gathering an index image (float32, exact): 0 = no source
    index_img = (torch.arange(hw, dtype=torch.float32) + 1).reshape(1, 1, h, w).expand(b, 1, h, w).contiguous().to(dev)
    src = RW.gather(index_img, theta, 0, None).reshape(b, hw).long().cpu() - 1
    assert (src[0] == src[0, 0]).all() and src[0, 0] >= 0 and (src[1] < 0).all()
    assert torch.equal(src[2], torch.arange(hw)) and src[3].unique(return_counts=True)[1].max() == 4
    g = (torch.randn(b, c, h, w, generator=torch.Generator().manual_seed(7)) * 2).to(dt)
    g[:, 0, 0, :8] = -0.0                               # a lone -0.0 contributor must come out as +0.0
    want = torch.zeros(b, c, hw, dtype=torch.float32)
    gf = g.float().reshape(b, c, hw)
    for i in range(b):
        ok = src[i] >= 0
        want[i].index_add_(1, src[i][ok], gf[i][:, ok])  # CPU index_add_: sequential, ascending output pixel
    want = want.to(dt).reshape(b, c, h, w)
    gd = g.to(dev)
    x = torch.zeros(b, c, h, w, dtype=dt, device=dev)
    plan = RW.build_inverse_plan(x, theta, 0, None)
    assert plan is not None
    got_plan = RW.gather_backward(gd, theta, 0, None, plan=plan)
    got_list = RW.gather_backward(gd, theta, 0, None)
    assert torch.equal(got_plan, got_list)
    if not half:
        assert torch.equal(got_plan.cpu(), want)
    else:
        # the sum of 4096 rounded-once values: same order, same float32 arithmetic -> same bits
        assert torch.equal(got_plan.cpu().float(), want.float())
    assert not torch.signbit(got_plan[2, 0, 0, :8]).any()
    del half


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("b,c", [(32, 16), (3, 21), (150, 5), (1, 64)])
def test_gather_decode_equals_gather_then_decode(dev, dt, b, c):
    """udape_rewarp_decode_select: the planes are arg-maxed where they are gathered, the re-warped map is never written.
    Every output of decode / decode_select on the materialised map — idx, preds, maxvals, position, conf_table, the k-th
    value threshold and tea_mask — bit for bit, on ordinary augmentations and on planes built to tie (constant planes,
    duplicated maxima, all-negative planes whose maximum is an out-of-image zero, NaN)."""
    from uda_poseestimation_b200 import keypoint_detection as KD
    y = S.heatmaps(b, c, seed=600 + b, peak=(0.3, 1.2))
    y[0, 0] = 0.25                                   # constant: ties everywhere -> first output pixel with a source
    y[0, 1 % c] = -1.0                               # all negative: the zeros of pixels that left the image win
    if c > 2:
        y[0, 2, 10, 11] = y[0, 2, 40, 41] = 3.0      # duplicated maximum
    if c > 3:
        y[0, 3, 5, 5] = float("nan")
    y = y.to(dt).to(dev)
    aug = S.aug_params(b, seed=610 + c)
    half = dt != torch.float32
    theta, half_mask, _ = RW.stage_table(RW.recon_stages(aug, 4.0, b), 64, 64, dt if half else torch.float32, dt if half else None)
    theta = theta.to(dev)
    grid = dt if half else None
    assert RW.gather_decode_supported(y)
    mid = RW.gather(y, theta, half_mask, grid)
    kth = max(1, (b * c) // 2)
    tm = (torch.rand(b, c, generator=torch.Generator().manual_seed(5)) > 0.2).float().to(dev)
    want = KD.decode(mid, want_idx=True, want_preds=True, want_maxvals_f32=True, want_position=True, occlude_thresh=0.9,
                     select_kth=kth, select_tea_mask=tm)
    got = RW.gather_decode(y, theta, half_mask, grid, want_idx=True, want_preds=True, want_position=True, occlude_thresh=0.9,
                           select_kth=kth, select_tea_mask=tm)
    for k in ("idx", "preds", "position", "conf_table", "tea_mask"):
        assert torch.equal(got[k], want[k]), k
    for k in ("maxvals_f32", "mask_thresh"):         # NaN-aware equality of the bits
        assert torch.equal(got[k].view(torch.int32), want[k].view(torch.int32)), k
    # no select, nothing optional
    lean = RW.gather_decode(y, theta, half_mask, grid, want_preds=False)
    assert set(lean) == {"maxvals_f32"} and torch.equal(lean["maxvals_f32"].view(torch.int32), want["maxvals_f32"].view(torch.int32))


def test_gather_decode_rejects_what_it_has_no_launch_for(dev):
    y = torch.zeros(2, 3, 32, 32, device=dev)
    theta = torch.zeros(2, 3, 6, device=dev)
    assert not RW.gather_decode_supported(y)
    with pytest.raises(ValueError, match="no fused launch"):
        RW.gather_decode(y, theta)
    import uda_poseestimation_b200 as U_
    tt = U_.teacher_targets_rewarped(y, theta, 2, 0.5, occlude_thresh=0.9)       # falls back to two launches
    assert tt["y_t_tea_recon"] is None and tt["tea_mask"].shape == (2, 3)

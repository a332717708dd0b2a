"""Host-side logic that runs without a GPU: PCK ratio formation from integer counts, the
coordinate helpers, sharding arithmetic, synthetic-input generators, and the multi-rank
(gloo, world_size 2) integer-count / flat-gradient all-reduce paths."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import reference_port as R
from uda_poseestimation_b200 import dist as D
from uda_poseestimation_b200 import keypoint_detection as KD
from uda_poseestimation_b200 import synthetic as S

ROOT = Path(__file__).resolve().parents[1]


def _pair(seed, b=6, k=5):
    joints, vis = S.keypoints(b, k, seed=seed)
    target = np.stack([R.generate_target(joints[i], vis[i], (64, 64), 2, (256, 256))[0] for i in range(b)])
    g = torch.Generator().manual_seed(seed)
    out = torch.roll(torch.from_numpy(target), shifts=(1, -2), dims=(2, 3)) + 0.02 * torch.randn(b, k, 64, 64, generator=g)
    return out.numpy(), target


def test_accuracy_from_counts_matches_reference_ratios():
    out, target = _pair(3)
    hits, valid, _ = R.pck_counts(out, target)
    acc_ref, avg_ref, cnt_ref, _ = R.accuracy(out, target)
    acc, avg, cnt = KD.accuracy_from_counts(hits, valid)
    np.testing.assert_array_equal(acc, acc_ref)
    assert avg == avg_ref and cnt == cnt_ref
    acc, avg, cnt = KD.accuracy_from_counts(np.zeros(4, np.int32), np.zeros(4, np.int32))
    assert (acc == -1).all() and avg == 0 and cnt == 0


def test_calc_dists_and_dist_acc_match_oracle():
    out, target = _pair(4)
    p, _ = R.get_max_preds(out)
    t, _ = R.get_max_preds(target)
    norm = np.ones((p.shape[0], 2)) * np.array([64, 64]) / 10
    d_ref = R.calc_dists(p, t, norm)
    d = KD.calc_dists(p, t, norm)
    np.testing.assert_array_equal(d, d_ref)
    for row in d:
        assert KD.dist_acc(row) == R.dist_acc(row)
    assert KD.dist_acc(np.full(5, -1.0)) == -1


def test_pck_is_the_integer_test_at_64():
    """at 64x64 / thr 0.5 the float64 distance test equals dx^2+dy^2 <= 10"""
    norm = np.ones((1, 2)) * np.array([64, 64]) / 10
    for dx in range(-8, 9):
        for dy in range(-8, 9):
            p = np.array([[[20.0 + dx, 20.0 + dy]]], dtype=np.float32)
            t = np.array([[[20.0, 20.0]]], dtype=np.float32)
            hit = KD.calc_dists(p, t, norm)[0, 0] < 0.5
            assert hit == (dx * dx + dy * dy <= 10)


def test_shard_bounds_cover_the_batch():
    for n in (1, 7, 32, 64, 255):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_generators():
    c, s = S.vgg_features(2, seed=1, channels=8)
    assert c.shape == (2, 8, 32, 32) and (c >= 0).all() and s.mean() > c.mean()
    assert torch.equal(c, S.vgg_features(2, seed=1, channels=8)[0])
    hm = S.heatmaps(3, 4, seed=2)
    assert hm.shape == (3, 4, 64, 64) and hm.amax(dim=(2, 3)).min() > 0.05
    shapes = S.pose_resnet_param_shapes(21)
    assert len(shapes) == 323 and sum(int(np.prod(x)) for x in shapes) == 52992853
    assert sum(int(np.prod(x)) for x in S.pose_resnet_param_shapes(16)) == 52991568
    adv = S.adversarial_heatmaps()
    assert torch.isnan(adv[3]).any() and (adv[2] < 0).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_np, tgt_np, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # PCK: local integer counts on this rank's shard (CPU oracle stands in for the CUDA kernel
    # here; the exchange logic is what is under test), int32 all-reduce, ratios afterwards
    o, t = D.shard(torch.from_numpy(out_np), rank, world).numpy(), D.shard(torch.from_numpy(tgt_np), rank, world).numpy()
    hits, valid, _ = R.pck_counts(o, t)
    counts = torch.from_numpy(np.stack([hits, valid]).astype(np.int32))
    D.allreduce_counts(counts)
    acc, avg, cnt = KD.accuracy_from_counts(counts[0], counts[1])
    # gradients: flat bucket, mean all-reduce
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    bucket = D.FlatGradBucket(model.parameters())
    x = torch.full((4, 5), float(rank + 1))
    model(x).sum().backward()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    bucket.allreduce_(average=True)
    loss = D.mean_scalar(torch.tensor(float(rank)))
    ret[rank] = dict(acc=acc, avg=avg, cnt=cnt, grad=bucket.flat.clone().numpy(), loss=float(loss))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_pck_and_gradient_allreduce():
    out, target = _pair(5, b=8, k=5)
    acc_ref, avg_ref, cnt_ref, _ = R.accuracy(out, target)
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, out, target, ret), nprocs=2, join=True)
    for rank in (0, 1):
        np.testing.assert_array_equal(ret[rank]["acc"], acc_ref)  # equals single-process PCK on the full batch
        assert ret[rank]["avg"] == avg_ref and ret[rank]["cnt"] == cnt_ref
        assert ret[rank]["loss"] == 0.5
    np.testing.assert_array_equal(ret[0]["grad"], ret[1]["grad"])
    # reference gradient: mean over the two ranks' inputs
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    gs = []
    for r in (0, 1):
        model.zero_grad()
        model(torch.full((4, 5), float(r + 1))).sum().backward()
        gs.append(torch.cat([p.grad.flatten() for p in model.parameters()]))
    expect = (gs[0] / 2 + gs[1] / 2).numpy()
    got = ret[0]["grad"]
    # bucket pads each tensor to a 16-byte boundary; compare the packed values
    sizes = [p.numel() for p in model.parameters()]
    off, vals = 0, []
    for n in sizes:
        off = (off + 3) // 4 * 4
        vals.append(got[off:off + n])
        off += n
    np.testing.assert_allclose(np.concatenate(vals), expect, rtol=1e-6, atol=1e-7)


# ---- re-warp host logic: the stage table the kernel consumes (no GPU) ------------------------------
def _emulate_table(theta, half_mask, grid_dtype, h, w):
    """numpy emulation of csrc/rewarp.cu::composed_source for one sample's [S,6] table."""
    f32 = np.float32

    def rnd(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(grid_dtype).float().numpy()

    jj, ii = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    i, j = ii.reshape(-1).astype(np.int64), jj.reshape(-1).astype(np.int64)
    alive = np.ones(h * w, dtype=bool)
    for s in range(theta.shape[0]):
        r = theta[s]
        half = (half_mask >> s) & 1
        x = (i.astype(f32) + f32(0.5 - 0.5 * w)).astype(f32)
        y = (j.astype(f32) + f32(0.5 - 0.5 * h)).astype(f32)
        if half:
            x, y = rnd(x), rnd(y)
        src = []
        for row, size in ((0, w), (1, h)):
            t = (x * r[3 * row]).astype(f32)
            t = (y.astype(np.float64) * np.float64(r[3 * row + 1]) + t.astype(np.float64)).astype(f32)
            g = (t + r[3 * row + 2]).astype(f32)
            if half:
                g = rnd(g)
            src.append(np.rint(((g + f32(1)) * f32(size) - f32(1)) * f32(0.5)))
        ok = (src[0] >= 0) & (src[0] <= w - 1) & (src[1] >= 0) & (src[1] <= h - 1)
        alive &= ok
        i = np.where(ok, src[0], 0).astype(np.int64)
        j = np.where(ok, src[1], 0).astype(np.int64)
    return np.where(alive, j * w + i, -1)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
def test_rewarp_stage_table_matches_oracle(dt):
    from oracle import reference_port as R
    from uda_poseestimation_b200 import rewarp as RW
    from uda_poseestimation_b200 import synthetic as S

    b, h, w, ratio = 6, 64, 64, 4.0
    aug = S.aug_params(b, seed=5, shear_y=True)
    ac = None if dt == torch.float32 else dt
    table, half_mask, code = RW.stage_table(RW.recon_stages(aug, ratio, b), h, w, dt, ac)
    assert table.shape == (b, 3, 6) and table.dtype == torch.float32
    assert half_mask == (0 if ac is None else 7)
    angle, [tx, ty], [sx, sy], sc = aug
    for i in range(b):
        want = R.recon_source_index(angle[i].item(), tx[i].item(), ty[i].item(), sx[i].item(), sy[i].item(),
                                    sc[i].item(), ratio, h, w, dt, ac)
        got = _emulate_table(table[i].numpy(), half_mask, dt, h, w)
        np.testing.assert_array_equal(got, want)


def test_rewarp_autocast_argument():
    from uda_poseestimation_b200 import rewarp as RW

    stages = [[(10.0, [1.0, 2.0], 1.1, [3.0, 0.0])]]
    with pytest.raises(NotImplementedError):
        RW.stage_table(stages, 8, 8, torch.float16, "auto")      # half tensor outside autocast
    with pytest.raises(NotImplementedError):
        RW.stage_table(stages, 8, 8, torch.float16, None)
    with pytest.raises(ValueError):
        RW.stage_table(stages, 8, 8, torch.float32, torch.float64)
    t32, m32, _ = RW.stage_table(stages, 8, 8, torch.float32, "auto")
    assert m32 == 0
    t16, m16, code = RW.stage_table(stages, 8, 8, torch.float32, torch.float16)  # fp32 tensor inside autocast(fp16)
    assert m16 == 1 and torch.equal(t16, t32.half().float())


def test_occlusion_plan_follows_reference_rng(golden):
    """The host plan draws from np.random in the reference's order (train_human.py:386-407)."""
    from oracle import reference_port as R
    from uda_poseestimation_b200 import rewarp as RW

    g = golden("rewarp")
    ratio, rate, size, image = g["occ_args"]
    a = torch.from_numpy(g["occ_aug"])
    aug = [a[:, 0], [a[:, 1].long(), a[:, 2].long()], [a[:, 3], a[:, 4]], a[:, 5]]
    active, paste, stages = RW.occlusion_plan(g["occ_conf_table"], g["occ_pred_position"], aug, float(ratio), float(rate),
                                              int(size), int(image), rng=np.random.RandomState(int(g["occ_seed"])))
    changed = (g["occ_in"] != g["occ_out"]).reshape(len(active), -1).any(1)
    np.testing.assert_array_equal(active.astype(bool), changed)
    assert all(len(s) == 4 for s in stages)
    # emulate the kernel on the host for the occluded samples: warp back, paste remap, three-stage warp
    table, half_mask, _ = RW.stage_table(stages, int(image), int(image), torch.float32, None)
    x_in, x_out = g["occ_in"], g["occ_out"]
    n = int(image)
    for bi in np.nonzero(active)[0]:
        th = table[bi].numpy()
        first = _emulate_table(th[:1], 0, torch.float32, n, n)
        j, i = first // n, first % n
        r0, r1, c0, c1, sr, scol = paste[bi]
        inside = (first >= 0) & (j >= r0) & (j < r1) & (i >= c0) & (i < c1)
        j = np.where(inside, j + sr - r0, j)
        i = np.where(inside, i + scol - c0, i)
        rest = _emulate_table(th[1:], 0, torch.float32, n, n)
        src = np.where(first >= 0, rest[np.clip(j * n + i, 0, None)], -1)
        flat = x_in[bi].reshape(3, -1)
        want = np.where(src >= 0, flat[:, np.clip(src, 0, None)], 0.0).reshape(3, n, n)
        np.testing.assert_array_equal(want, x_out[bi])


def test_bench_reference_arm_prints_one_json_line():
    """Driver contract: stdout of bench.py is ONE JSON line (fd 1 is re-pointed at stderr for everything
    else); the reference arm runs the CPU path and needs no GPU."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hot_path_images_per_sec" and d["unit"] == "images/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    # "reference": the reference's own functions ran (from /root/reference here, from oracle/_ref bytecode on the GPU
    # box); "port": neither was present
    from oracle import ref_loader
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_numa_binding_is_best_effort():
    from uda_poseestimation_b200 import dist as D
    import os
    before = os.sched_getaffinity(0)
    cpus = D.bind_to_gpu_numa(0)          # no NVML / no GPU here: must not raise and must not shrink the mask to nothing
    assert cpus is None or len(cpus) >= 1
    assert len(os.sched_getaffinity(0)) >= 1
    os.sched_setaffinity(0, before)


def test_old_weight_ema_is_constructible_before_the_models_move_to_the_gpu():
    """train_human.py:141 builds OldWeightEMA(teacher, student) BEFORE DataParallel(...).cuda() (:145-146): the
    constructor's initial copy (utils.py:18-19) must work on CPU modules; only step() is the CUDA operator."""
    import uda_poseestimation_b200 as U
    torch.manual_seed(0)
    student, teacher = torch.nn.Linear(7, 5), torch.nn.Linear(7, 5)
    ema = U.OldWeightEMA(teacher, student, alpha=0.999)
    for t, s in zip(teacher.parameters(), student.parameters()):
        assert torch.equal(t.data, s.data) and t.data_ptr() != s.data_ptr()
    assert ema.target_params[0] is next(teacher.parameters())      # the live Parameter objects (they survive .cuda())
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ema.step()                                                  # no CPU fallback for the operator itself


def test_step_bytes_with_the_teacher_chain_in_one_launch():
    """hotpath.step_algorithmic_bytes: with the teacher re-warp arg-maxed where it is gathered the step loses the map's
    write and its re-read by the decode (2 x B*K*H*W*4 bytes); k > 1 views or the unfused losses keep both launches."""
    from uda_poseestimation_b200.hotpath import StepInputs, step_algorithmic_bytes
    b, k = 4, 3
    hm32 = lambda: torch.zeros(b, k, 64, 64)
    f = torch.zeros(2, 8, 32, 32)
    one = torch.zeros(1)
    inp = StepInputs(f, f, f, f, hm32().half(), hm32().half(), hm32(), hm32(), torch.ones(b, k, 1), one, one,
                     theta_tea=torch.zeros(b, 3, 6), theta_stu=torch.zeros(b, 3, 6))
    two = step_algorithmic_bytes(inp, 1000, fused=True)
    fused = step_algorithmic_bytes(inp, 1000, fused=True, fuse_teacher_decode=True)
    assert two["total"] - fused["total"] == 2 * b * k * 64 * 64 * 4
    assert "decode" not in fused and "rewarp_teacher" not in fused and "rewarp_teacher+decode" in fused
    assert step_algorithmic_bytes(inp, 1000, fused=False, fuse_teacher_decode=True) == step_algorithmic_bytes(inp, 1000, fused=False)
    inp.y_t_tea, inp.theta_tea = [hm32(), hm32()], [torch.zeros(b, 3, 6)] * 2
    assert step_algorithmic_bytes(inp, 1000, fuse_teacher_decode=True) == step_algorithmic_bytes(inp, 1000)


def test_gather_decode_support_predicate_and_argument_errors():
    """rewarp.gather_decode_supported is pure host logic; the fused entry point rejects what it has no launch for
    through the C-ABI's error codes (no GPU needed: the checks run before any launch)."""
    import ctypes
    from uda_poseestimation_b200 import _lib, rewarp as RW
    assert not RW.gather_decode_supported(torch.zeros(2, 3, 64, 64))            # CPU tensor
    assert not RW.gather_decode_supported(torch.zeros(2, 3, 64))                 # not [B,C,H,W]
    lib = _lib.load()
    buf = (ctypes.c_float * 64)()
    addr = (ctypes.addressof(buf) + 15) & ~15                                      # a 16-byte aligned (host) address: never dereferenced
    args = dict(stages=3, half_mask=0, grid=_lib.F16, B=2, C=3, H=64, W=64, dtype=_lib.F32)
    def call(**kw):
        a = dict(args, **kw)
        return lib.udape_rewarp_decode_select(addr, addr, a["stages"], a["half_mask"], a["grid"], a["B"], a["C"], a["H"], a["W"],
                                              a["dtype"], None, None, addr, None, 0.9, None, a.get("kth", 0), None, None, None,
                                              a.get("ticket"), None)
    assert call(H=32, W=32) < 0                                                   # planes of 1024 pixels: UDAPE_ERR_SHAPE
    assert call(C=65) < 0                                                         # more planes than the maxima table holds
    assert call(kth=7) < 0 and call(kth=3, ticket=None) < 0                       # kth outside [0, B*C]; a select without a ticket
    assert call(stages=5) < 0
    msg = ctypes.create_string_buffer(256)
    lib.udape_last_error(msg, 256)
    assert b"stages" in msg.value

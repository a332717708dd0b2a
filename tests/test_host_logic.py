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

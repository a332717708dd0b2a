"""Re-entrancy (SURVEY.md §8b "Threading / streams"): ``nn.DataParallel.parallel_apply`` calls the style net — hence
``adain`` — from several Python threads at once, one per device (train_human.py:145-146, Style_net.py:163-168).  The
library keeps no global state but a thread-local error string and launches on the CALLING thread's current device
and stream; the host side hands every launch its own ticket word.  Here: many threads, each on its own stream (and
on its own device where the box has several), run a mix of operators concurrently and must reproduce, bit for bit,
what one thread computes alone; an argument error raised in one thread must not leak into another's.
"""
import threading

import numpy as np
import pytest
import torch

import uda_poseestimation_b200 as U
from uda_poseestimation_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _work(dev, seed):
    """One thread's share of a step: both AdaIN directions, teacher decode + masks, both losses with backward, PCK."""
    c, s = S.vgg_features(4, seed=seed, channels=64)
    tea = S.heatmaps(8, 16, seed=seed + 1, peak=(0.3, 1.2)).to(dev)
    y = S.heatmaps(8, 16, seed=seed + 2).half().to(dev)
    joints, vis = S.keypoints(8, 16, seed=seed + 3)
    label, weight = U.generate_target_batched(joints, vis, (64, 64), 2, (256, 256), device=dev)
    out = {"adain": U.adain_mix(c.to(dev), s.to(dev), 0.37)}
    t = U.teacher_targets(tea, 2, 0.5, occlude_thresh=0.9)
    out.update(tea_mask=t["tea_mask"], conf=t["conf_table"], rect=t["rectified"], thresh=t["mask_thresh"])
    o = y.clone().requires_grad_(True)
    loss = U.JointsMSELoss()(o, label, weight) + U.ConsLoss()(o, t["rectified"], tea_mask=t["tea_mask"])
    (loss * 65536.0).backward()
    out.update(loss=loss.detach(), grad=o.grad)
    hits, valid, pred = U.pck_counts(y, label)
    out.update(hits=hits, valid=valid, pred=pred)
    return {k: (v.detach().cpu() if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))) for k, v in out.items()}


def _run_threads(devices, rounds=6):
    want = {}
    for i, dev in enumerate(devices):
        with torch.cuda.device(dev):
            want[i] = _work(dev, 100 + i)
            torch.cuda.synchronize(dev)
    got, errors = {}, []
    start = threading.Barrier(len(devices))

    def body(i, dev):
        try:
            with torch.cuda.device(dev), torch.cuda.stream(torch.cuda.Stream(dev)):
                start.wait()
                for _ in range(rounds):
                    r = _work(dev, 100 + i)
                    # an argument error in this thread: its message stays this thread's
                    with pytest.raises((ValueError, TypeError, AssertionError, RuntimeError)):   # (UdapeError is a RuntimeError)
                        U.adain_mix(torch.zeros(2, 3, 4, 4, device=dev), torch.zeros(2, 5, 4, 4, device=dev), 0.5)
                got[i] = r
        except BaseException as exc:  # noqa: BLE001 — reported by the main thread
            errors.append((i, repr(exc)))

    threads = [threading.Thread(target=body, args=(i, d)) for i, d in enumerate(devices)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(300)
    assert not errors, errors
    for i in want:
        for k, v in want[i].items():
            assert torch.equal(got[i][k], v), (i, k)


def test_eight_threads_on_one_device(dev):
    _run_threads([dev] * 8)
    U.check_tickets()      # every ticket word is back at zero


def test_one_thread_per_device():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (nn.DataParallel's one thread per device)")
    _run_threads([torch.device("cuda", i) for i in range(n)])
    U.check_tickets()

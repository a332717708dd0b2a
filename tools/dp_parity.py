#!/usr/bin/env python
"""On-hardware N-rank parity of the two exchanges of the data-parallel step (SURVEY.md §4 tier 5):

* the student-gradient exchange + update: after ``ShardedStudentStep.step()`` on N real ranks (one process per
  GPU, peer loads over NVLink) every rank's student AND teacher equal — bit for bit — the single-process fused
  step (``udape_student_step``) applied to the rank-ordered mean of every rank's seeded gradient; the NCCL
  flat-bucket all-reduce equals the same mean to 1e-6 relative (NCCL's ring order differs per chunk);
* the integer PCK exchange: the all-reduced ``hits || valid`` of the ranks' shards equal ``pck_counts`` of the
  concatenated batch computed on one GPU, through both the peer-memory kernel and ``ncclAllReduce``.

Run under torchrun (``python -m torch.distributed.run --nproc-per-node N tools/dp_parity.py``); rank 0 prints
``multi_gpu_parity ok``.  ``check()`` is what ``bench.py`` calls for the ``multi_gpu_parity`` key of its line.
Every rank regenerates every other rank's inputs from the seeds (``1234 + rank``), so nothing but the results
under test crosses the wire.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import uda_poseestimation_b200 as U  # noqa: E402
from uda_poseestimation_b200 import dist as D  # noqa: E402
from uda_poseestimation_b200 import dp as DP  # noqa: E402
from uda_poseestimation_b200 import synthetic as S  # noqa: E402


class _Bag(torch.nn.Module):
    def __init__(self, tensors):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in tensors])


def _cat(ts):
    return torch.cat([t.detach().reshape(-1) for t in ts])


def _rank_grads(shapes, rank: int, it: int, dev, world: int):
    g = torch.Generator(device=dev).manual_seed(100_000 * (it + 1) + 1234 + rank)
    out = [torch.randn(s, generator=g, device=dev) * 40.0 for s in shapes]
    if it == 1 and rank == world - 1:
        out[len(out) // 2].view(-1)[3] = float("inf")      # one rank overflows: the step is skipped EVERYWHERE
    return out


def check(dev: torch.device, keypoints: int = 16, batch: int = 32, steps: int = 3, algo: str = "adam") -> dict:
    """Collective over the default process group.  Returns a dict of booleans (identical on every rank)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    shapes = S.pose_resnet_param_shapes(keypoints)
    init = S.parameter_list(shapes, seed=77, device=dev)
    _, n_total = DP.flat_layout(init)
    peers = U.PeerGroup.create(DP.arena_bytes(n_total), dev)
    stu, tea = _Bag(init), _Bag(init)
    kw = dict(lr=1e-3) if algo == "adam" else dict(lr=0.05, momentum=0.9, weight_decay=1e-4, nesterov=True)
    opt = U.ShardedStudentStep(stu.parameters(), peers, algo=algo, teacher_params=list(tea.parameters()), alpha=0.999, **kw)
    ref_s, ref_t = _Bag(init), _Bag(init)
    ref_opt = (U.Adam if algo == "adam" else U.SGD)(ref_s.parameters(), **kw)
    ref_tea = U.OldWeightEMA(ref_t, ref_s, alpha=0.999)
    ref_opt.attach_teacher(ref_tea)
    bucket = D.FlatGradBucket(list(_Bag(init).parameters()))
    scale = torch.full((), 1024.0, device=dev)
    ok = dict(student=True, teacher=True, found_inf=True, nccl_allreduce=True, pck_peer=True, pck_nccl=True)
    inv = torch.tensor(1.0 / world, dtype=torch.float32, device=dev)
    for it in range(steps):
        everyone = [_rank_grads(shapes, q, it, dev, world) for q in range(world)]
        for p, g in zip(stu.parameters(), everyone[rank]):
            p.grad.copy_(g)
        opt.grad_scale = scale
        opt.step()
        torch.cuda.synchronize()
        opt.check()
        mean = []
        for i in range(len(shapes)):
            acc = everyone[0][i].clone()
            for q in range(1, world):
                acc = acc + everyone[q][i]
            mean.append(acc * inv)
        for p, g in zip(ref_s.parameters(), mean):
            p.grad = g
        ref_opt.grad_scale, ref_opt.found_inf = scale, ref_opt.check_grads()
        ref_opt.step()
        ref_tea.step()
        ok["found_inf"] &= float(opt.found_inf) == float(ref_opt.found_inf) == (1.0 if it == 1 else 0.0)
        ok["student"] &= torch.equal(_cat(stu.parameters()), _cat(ref_s.parameters()))
        ok["teacher"] &= torch.equal(_cat(tea.parameters()), _cat(ref_t.parameters()))
        if it == 0:
            # NCCL path: flat bucket all-reduce (mean) against the same single-process sum
            for v, g in zip(bucket.views, everyone[rank]):
                v.copy_(g)
            bucket.allreduce_(average=True)
            want = _cat(mean)
            got = torch.cat([v.reshape(-1) for v in bucket.views])
            err = (got - want).abs().max() / want.abs().max()
            ok["nccl_allreduce"] &= bool(err <= 1e-6)
    # PCK counts: shards -> exchange -> == counts of the concatenated batch
    hm = [S.heatmaps(batch, keypoints, 1234 + q + 2, peak=(0.2, 1.1)).to(torch.float16).to(dev) for q in range(world)]
    lab = []
    for q in range(world):
        j, v = S.keypoints(batch, keypoints, 1234 + q + 5)
        lab.append(U.generate_target_batched(torch.from_numpy(j).to(dev), torch.from_numpy(v).to(dev), (64, 64), 2, (256, 256), device=dev)[0])
    from uda_poseestimation_b200.keypoint_detection import _pck
    whole, _ = _pck(torch.cat(hm), torch.cat(lab), 0.5)      # int32 [2,K] = hits || valid
    mine, _ = _pck(hm[rank], lab[rank], 0.5)
    got = opt.allreduce_counts(mine)
    torch.cuda.synchronize()
    opt.check()
    ok["pck_peer"] &= torch.equal(got, whole)
    nccl = mine.clone()
    D.allreduce_counts(nccl)
    ok["pck_nccl"] &= torch.equal(nccl, whole)
    flags = torch.tensor([int(v) for v in ok.values()], dtype=torch.int32, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    peers_ok = {k: bool(f) for k, f in zip(ok, flags.tolist())}
    del opt, stu, tea
    peers.close()
    return peers_ok


def main():
    rank, world, local = D.init_from_env("nccl")
    if world < 2:
        raise SystemExit("dp_parity.py: run under torchrun with >= 2 ranks")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    res = check(dev)
    if rank == 0:
        print(("multi_gpu_parity ok " if all(res.values()) else "multi_gpu_parity FAILED ") + str(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not all(res.values()):
        raise SystemExit(1)


if __name__ == "__main__":
    main()

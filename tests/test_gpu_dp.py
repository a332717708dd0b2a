"""The data-parallel tail of the step over peer memory (csrc/dp.cu, uda_poseestimation_b200/dp.py).

Parity bar (SURVEY.md §4 tier 5): the N-rank result equals the single-process result on the rank-summed
gradient — BIT-EXACT here, because the reduce-scatter sums in rank order and the sharded update runs the same
element arithmetic as udape_student_step; the integer PCK-count exchange equals the plain sum.  On a one-GPU box
the N ranks are `PeerGroup.virtual` ranks (N arenas on one device, one stream per rank, the same kernels and
signalling protocol); with >= 2 GPUs the same checks run as real processes over CUDA IPC (test at the bottom).
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

import uda_poseestimation_b200 as U
from conftest import assert_close_scaled
from oracle import reference_port as R
from uda_poseestimation_b200 import dp as DP
from uda_poseestimation_b200 import synthetic as S

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


class Bag(torch.nn.Module):
    def __init__(self, tensors):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in tensors])


def _cat(ts):
    return torch.cat([t.detach().float().reshape(-1) for t in ts])


def _mean_in_rank_order(grads):
    """what the reduce-scatter computes: ((g0 + g1) + g2) + ... then * (1/W), all float32"""
    acc = grads[0].clone()
    for g in grads[1:]:
        acc = acc + g
    return acc * torch.tensor(1.0 / len(grads), dtype=torch.float32, device=acc.device)


def _run_ranks(ranks, fn):
    """one stream per virtual rank: the ranks' kernels must be able to run concurrently (they wait for each other)"""
    streams = [torch.cuda.Stream() for _ in ranks]
    cur = torch.cuda.current_stream()
    for st in streams:
        st.wait_stream(cur)
    for r, st in zip(ranks, streams):
        with torch.cuda.stream(st):
            fn(r)
    for st in streams:
        cur.wait_stream(st)
    torch.cuda.synchronize()


SHAPES = [(300,), (17, 9), (5000,), (64, 3, 7, 7), (1,), (4099,), (256, 33)]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("algo", ["adam", "sgd"])
def test_sharded_step_equals_the_replicated_step(dev, world, algo):
    """3 scaled steps (the 2nd with an inf in ONE rank's bucket: skipped everywhere, EMA still applied) on W
    virtual ranks vs the single-GPU fused step (udape_student_step) fed the rank-ordered mean gradient."""
    torch.manual_seed(world)
    cpu = [torch.randn(s) * 0.1 for s in SHAPES]
    kw = dict(lr=1e-2) if algo == "adam" else dict(lr=0.05, momentum=0.9, weight_decay=1e-4, nesterov=True)
    # reference: one process, replicated update
    ref_s, ref_t = Bag(cpu).to(dev), Bag(cpu).to(dev)
    ref_opt = (U.Adam if algo == "adam" else U.SGD)(ref_s.parameters(), **kw)
    ref_tea = U.OldWeightEMA(ref_t, ref_s, alpha=0.99)
    ref_opt.attach_teacher(ref_tea)
    # W ranks
    _, n_total = DP.flat_layout(cpu)
    groups = U.PeerGroup.virtual(world, DP.arena_bytes(n_total), dev)
    stu = [Bag(cpu).to(dev) for _ in range(world)]
    tea = [Bag(cpu).to(dev) for _ in range(world)]
    opts = [U.ShardedStudentStep(stu[r].parameters(), groups[r], algo=algo, teacher_params=list(tea[r].parameters()),
                                 alpha=0.99, timeout_s=5.0, **kw) for r in range(world)]
    scale = torch.full((), 1024.0, device=dev)
    g = torch.Generator().manual_seed(5)
    for it in range(3):
        per_rank = [[torch.randn(t.shape, generator=g) * 10.0 for t in cpu] for _ in range(world)]
        if it == 1:
            per_rank[world - 1][2].view(-1)[7] = float("inf")
        for r in range(world):
            for p, gr in zip(stu[r].parameters(), per_rank[r]):
                p.grad.copy_(gr.to(dev))          # p.grad IS the peer-readable bucket
        for o in opts:
            o.grad_scale = scale
        _run_ranks(range(world), lambda r: opts[r].step())
        for o in opts:
            o.check()
            assert float(o.found_inf) == (1.0 if it == 1 else 0.0)
        mean = [_mean_in_rank_order([per_rank[r][i].to(dev) for r in range(world)]) for i in range(len(cpu))]
        for p, gr in zip(ref_s.parameters(), mean):
            p.grad = gr
        ref_opt.grad_scale, ref_opt.found_inf = scale, ref_opt.check_grads()
        assert float(ref_opt.found_inf) == (1.0 if it == 1 else 0.0)
        ref_opt.step()
        ref_tea.step()
        for r in range(world):
            assert torch.equal(_cat(stu[r].parameters()), _cat(ref_s.parameters())), f"student, rank {r}, step {it}"
            assert torch.equal(_cat(tea[r].parameters()), _cat(ref_t.parameters())), f"teacher, rank {r}, step {it}"
        # the gradient buckets are inputs: the step leaves them as backward wrote them
        for r in range(world):
            for p, gr in zip(stu[r].parameters(), per_rank[r]):
                assert torch.equal(p.grad, gr.to(dev)) or it == 1
    assert all(o.applied_steps() == 2 for o in opts) and ref_opt.applied_steps() == 2
    # the sharded optimizer state, put back together, is the replicated optimizer's state
    key1 = "exp_avg" if algo == "adam" else "momentum_buffer"
    flat1 = torch.zeros(n_total, device=dev)
    flat2 = torch.zeros(n_total, device=dev)
    for off, p in zip(opts[0].offsets, ref_s.parameters()):
        flat1[off:off + p.numel()] = ref_opt.state[p][key1].reshape(-1)
        if algo == "adam":
            flat2[off:off + p.numel()] = ref_opt.state[p]["exp_avg_sq"].reshape(-1)
    for r in range(world):
        lo, hi = opts[r].shard_bounds()
        s1, s2 = opts[r].state_shards()
        assert torch.equal(s1[:hi - lo], flat1[lo:hi]), f"state1 shard of rank {r}"
        if algo == "adam":
            assert torch.equal(s2[:hi - lo], flat2[lo:hi]), f"state2 shard of rank {r}"
    for grp in groups:
        grp.close()


def test_sharded_step_pose_resnet101_census_vs_oracle(dev):
    """The full 323-tensor PoseResNet-101 parameter list on 2 virtual ranks, two Adam steps, against the CPU
    oracle (torch.optim.Adam + the reference's OldWeightEMA on the rank-mean gradient): 1e-5 relative."""
    shapes = S.pose_resnet_param_shapes(16)
    cpu = S.parameter_list(shapes, seed=3)
    world = 2
    _, n_total = DP.flat_layout(cpu)
    groups = U.PeerGroup.virtual(world, DP.arena_bytes(n_total), dev)
    stu = [Bag(cpu).to(dev) for _ in range(world)]
    tea = [Bag(cpu).to(dev) for _ in range(world)]
    opts = [U.ShardedStudentStep(stu[r].parameters(), groups[r], algo="adam", teacher_params=list(tea[r].parameters()),
                                 alpha=0.999, lr=1e-3, timeout_s=10.0) for r in range(world)]
    s_cpu, t_cpu = [t.clone() for t in cpu], [t.clone() for t in cpu]
    m_cpu, v_cpu = [torch.zeros_like(t) for t in cpu], [torch.zeros_like(t) for t in cpu]
    g = torch.Generator().manual_seed(11)
    for it in range(2):
        per_rank = [[torch.randn(t.shape, generator=g) * 0.01 for t in cpu] for _ in range(world)]
        for r in range(world):
            for p, gr in zip(stu[r].parameters(), per_rank[r]):
                p.grad.copy_(gr.to(dev))
        _run_ranks(range(world), lambda r: opts[r].step())
        for o in opts:
            o.check()
        mean = [(a + b) * 0.5 for a, b in zip(*per_rank)]
        R.adam_step(s_cpu, mean, m_cpu, v_cpu, it + 1, 1e-3)
        R.ema_step(t_cpu, s_cpu, 0.999)
    for r in range(world):
        assert_close_scaled(_cat(stu[r].parameters()).cpu(), _cat(s_cpu), 1e-5, f"student rank {r}")
        assert_close_scaled(_cat(tea[r].parameters()).cpu(), _cat(t_cpu), 1e-5, f"teacher rank {r}")
    assert torch.equal(_cat(stu[0].parameters()), _cat(stu[1].parameters()))
    for grp in groups:
        grp.close()


@pytest.mark.parametrize("world", [2, 8])
def test_pck_count_exchange_equals_the_sum(dev, world):
    """int32 hits || valid of every rank summed in one single-CTA launch per rank; repeated so that both parities
    of the double-buffered slots and the monotonic step numbers are exercised.  Exact."""
    k = 21
    groups = U.PeerGroup.virtual(world, DP.arena_bytes(64), dev)
    params = [[torch.nn.Parameter(torch.zeros(64, device=dev))] for _ in range(world)]
    opts = [U.ShardedStudentStep(params[r], groups[r], timeout_s=5.0) for r in range(world)]
    g = torch.Generator().manual_seed(1)
    for it in range(5):
        counts = [torch.randint(0, 300, (2, k), generator=g, dtype=torch.int32).to(dev) for _ in range(world)]
        outs = [None] * world

        def go(r):
            outs[r] = opts[r].allreduce_counts(counts[r])
        _run_ranks(range(world), go)
        want = torch.stack(counts).sum(0).to(torch.int32)
        for r in range(world):
            opts[r].check()
            assert torch.equal(outs[r], want), f"rank {r}, round {it}"
    for grp in groups:
        grp.close()


def test_a_missing_rank_times_out_instead_of_hanging(dev):
    """Only rank 0 of a 2-rank group steps: the bounded wait sets the error word and check() raises."""
    groups = U.PeerGroup.virtual(2, DP.arena_bytes(64), dev)
    p = [torch.nn.Parameter(torch.zeros(64, device=dev))]
    opt = U.ShardedStudentStep(p, groups[0], timeout_s=0.2)
    opt.step()
    torch.cuda.synchronize()
    with pytest.raises(U.UdapeError, match="timed out"):
        opt.check()
    for grp in groups:
        grp.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_two_processes_over_cuda_ipc():
    """Real ranks: tools/dp_parity.py under torchrun (one process per GPU, arenas mapped through CUDA IPC, peer
    loads over NVLink) prints 'multi_gpu_parity ok' after comparing with the single-process result."""
    n = min(torch.cuda.device_count(), 8)
    n = 1 << (n.bit_length() - 1)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(ROOT / "tools" / "dp_parity.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multi_gpu_parity ok" in r.stdout

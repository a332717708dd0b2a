"""GPU parity of the fused student step (unscale + Adam | SGD + teacher EMA, csrc/optim.cu) against the
CPU oracle (oracle/reference_port.py::student_teacher_step, pinned to torch.optim + the reference's
OldWeightEMA + torch.amp.GradScaler by tests/golden/optim.npz) and against torch's own optimizers run on
the same device.  Bar: 1e-5 relative, scale-aware (north star: "EMA weights in fp32").
"""
import numpy as np
import pytest
import torch

import uda_poseestimation_b200 as U
from conftest import assert_close_scaled
from oracle import reference_port as R
from uda_poseestimation_b200 import synthetic as S

pytestmark = pytest.mark.gpu
RTOL = 1e-5


class Bag(torch.nn.Module):
    def __init__(self, tensors):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in tensors])


def _cat(ts):
    return torch.cat([t.detach().float().cpu().reshape(-1) for t in ts])


@pytest.mark.parametrize("algo", ["adam", "sgd"])
def test_student_step_pose_resnet101_vs_oracle(dev, algo):
    """The full 323-tensor / 52 992 853-parameter PoseResNet-101 census (K=21), 3 scaled steps with the
    second one carrying a NaN gradient (skipped update, EMA still applied)."""
    shapes = S.pose_resnet_param_shapes(21)
    student_cpu = S.parameter_list(shapes, seed=1)
    grads_seed = 7
    student = Bag(student_cpu).to(dev)
    teacher = Bag(student_cpu).to(dev)
    teacher_cpu = [t.clone() for t in student_cpu]
    if algo == "adam":
        opt = U.Adam(student.parameters(), lr=1e-3)
        hyper = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    else:
        opt = U.SGD(student.parameters(), lr=0.1, momentum=0.9, weight_decay=0.0001, nesterov=True)
        hyper = dict(lr=0.1, momentum=0.9, dampening=0.0, weight_decay=0.0001, nesterov=True)
    tea = U.OldWeightEMA(teacher, student, alpha=0.999)
    opt.attach_teacher(tea)
    s1 = [torch.zeros_like(t) for t in student_cpu]
    s2 = [torch.zeros_like(t) for t in student_cpu]
    scale = torch.full((), 65536.0, device=dev)
    step = 0
    g = torch.Generator().manual_seed(grads_seed)
    for it in range(3):
        grads = [torch.randn(t.shape, generator=g) * (0.01 * 65536.0) for t in student_cpu]
        if it == 1:
            grads[200].view(-1)[5] = float("nan")
        for p, gr in zip(student.parameters(), grads):
            p.grad = gr.to(dev)
        opt.grad_scale, opt.found_inf = scale, opt.check_grads()
        assert float(opt.found_inf) == (1.0 if it == 1 else 0.0)
        opt.step()
        tea.step()   # folded into opt.step(): must be a no-op
        found, step = R.student_teacher_step(algo, student_cpu, grads, s1, s2, teacher_cpu, step, 65536.0, 0.999, **hyper)
        assert found == (it == 1)
        assert_close_scaled(_cat(student.parameters()), _cat(student_cpu), RTOL, f"student step {it}")
        assert_close_scaled(_cat(teacher.parameters()), _cat(teacher_cpu), RTOL, f"teacher step {it}")
    assert opt.applied_steps() == 2
    del opt.grad_scale, opt.found_inf


def test_student_step_ragged_frozen_and_missing_grads(dev):
    """Odd sizes / unaligned tails, a parameter without gradient (EMA only), a frozen student parameter the
    optimizer does not own (EMA only through attach_teacher), two param groups, no GradScaler."""
    torch.manual_seed(3)
    shapes = [(7,), (1,), (33, 5), (4096,), (4099,), (2, 3, 5, 7), (18,), (129, 65)]
    cpu = [torch.randn(s) for s in shapes]
    student, teacher = Bag(cpu).to(dev), Bag(cpu).to(dev)
    ps = list(student.parameters())
    ps[6].requires_grad_(False)                      # frozen: not handed to the optimizer
    opt = U.Adam([{"params": ps[:3]}, {"params": ps[3:6] + ps[7:], "lr": 5e-3, "weight_decay": 0.1}], lr=1e-3)
    tea = U.OldWeightEMA(teacher, student, alpha=0.9)
    opt.attach_teacher(tea)
    s_cpu = [t.clone() for t in cpu]
    t_cpu = [t.clone() for t in cpu]
    m = [torch.zeros_like(t) for t in cpu]
    v = [torch.zeros_like(t) for t in cpu]
    for it in range(3):
        grads = [torch.randn(s) for s in shapes]
        grads[1] = None                                 # never receives a gradient
        grads[6] = None
        for p, gr in zip(ps, grads):
            p.grad = None if gr is None else gr.to(dev)
        opt.step()
        tea.step()
        with torch.no_grad():
            ga = [grads[i] for i in (0, 1, 2)]
            R.adam_step(s_cpu[:3], ga, m[:3], v[:3], it + 1, 1e-3, (0.9, 0.999), 1e-8, 0.0)
            idx = [3, 4, 5, 7]
            R.adam_step([s_cpu[i] for i in idx], [grads[i] for i in idx], [m[i] for i in idx], [v[i] for i in idx],
                        it + 1, 5e-3, (0.9, 0.999), 1e-8, 0.1)
            R.ema_step(t_cpu, s_cpu, 0.9)
        for i, (p, ref) in enumerate(zip(student.parameters(), s_cpu)):
            assert_close_scaled(p, ref, RTOL, f"student[{i}] step {it}")
        for i, (p, ref) in enumerate(zip(teacher.parameters(), t_cpu)):
            assert_close_scaled(p, ref, RTOL, f"teacher[{i}] step {it}")
    assert torch.equal(ps[1].detach().cpu(), cpu[1]) and torch.equal(ps[6].detach().cpu(), cpu[6])


def test_unfused_ema_order_and_plain_optimizer_use(dev):
    """Without attach_teacher the optimizer is a plain torch.optim drop-in and OldWeightEMA.step runs its
    own launch; both orders give the same teacher as the fused launch."""
    torch.manual_seed(5)
    cpu = [torch.randn(1000), torch.randn(64, 33)]
    outs = []
    for fuse in (False, True):
        student, teacher = Bag(cpu).to(dev), Bag(cpu).to(dev)
        opt = U.SGD(student.parameters(), lr=0.05, momentum=0.9)
        tea = U.OldWeightEMA(teacher, student, alpha=0.95)
        if fuse:
            opt.attach_teacher(tea)
        g = torch.Generator().manual_seed(1)
        for _ in range(4):
            for p in student.parameters():
                p.grad = torch.randn(p.shape, generator=g).to(dev)
            opt.step()
            tea.step()
        outs.append((_cat(student.parameters()), _cat(teacher.parameters())))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("algo", ["adam", "sgd"])
def test_amp_training_loop_vs_torch(dev, algo):
    """A real autocast + GradScaler loop (train_human.py:414-440 in miniature): the drop-in classes against
    torch.optim + torch.amp.GradScaler + the oracle EMA on the same device, same data, 6 steps."""
    def make():
        torch.manual_seed(11)
        return torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(16, 4, 1)).to(dev)

    ref_s, ref_t, new_s, new_t = make(), make(), make(), make()
    if algo == "adam":
        ref_opt, new_opt = torch.optim.Adam(ref_s.parameters(), lr=1e-3), U.Adam(new_s.parameters(), lr=1e-3)
    else:
        kw = dict(lr=0.05, momentum=0.9, weight_decay=1e-4, nesterov=True)
        ref_opt, new_opt = torch.optim.SGD(ref_s.parameters(), **kw), U.SGD(new_s.parameters(), **kw)
    new_tea = U.OldWeightEMA(new_t, new_s, alpha=0.99)
    new_opt.attach_teacher(new_tea)
    R.ema_init(list(ref_t.parameters()), list(ref_s.parameters()))
    ref_scaler, new_scaler = torch.amp.GradScaler("cuda", init_scale=1024.0), U.GradScaler(init_scale=1024.0)
    sched_ref = torch.optim.lr_scheduler.MultiStepLR(ref_opt, [3], 0.1)
    sched_new = torch.optim.lr_scheduler.MultiStepLR(new_opt, [3], 0.1)   # train_human.py:143
    g = torch.Generator().manual_seed(2)
    crit = U.JointsMSELoss()
    skipped = 0
    for it in range(6):
        x = torch.randn(4, 3, 16, 16, generator=g).to(dev)
        y = torch.randn(4, 4, 16, 16, generator=g).to(dev)
        w = torch.ones(4, 4, 1, device=dev)
        for model, opt, scaler, tea in ((ref_s, ref_opt, ref_scaler, None), (new_s, new_opt, new_scaler, new_tea)):
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.float16):
                loss = crit(model(x), y, w)
            if it == 2:
                loss = loss * float("inf")       # an overflowing step: every gradient is inf / nan
            scaler.scale(loss).backward()
            scaler.step(opt)
            if tea is None:
                with torch.no_grad():
                    R.ema_step(list(ref_t.parameters()), list(ref_s.parameters()), 0.99)
            else:
                tea.step()
            scaler.update()
        sched_ref.step()
        sched_new.step()
        assert ref_scaler.get_scale() == new_scaler.get_scale() == (1024.0 if it < 2 else 512.0)
    assert new_opt.applied_steps() == 5
    # fp16 forward/backward: 1e-7 differences in the weights flip fp16 roundings of activations, so the two
    # runs drift by a few 1e-6 relative per step; the fp32 update itself is held to 1e-5 by the tests above
    assert_close_scaled(_cat(new_s.parameters()), _cat(ref_s.parameters()), 1e-4, "student (fp16 autocast loop)")
    assert_close_scaled(_cat(new_t.parameters()), _cat(ref_t.parameters()), 1e-4, "teacher (fp16 autocast loop)")


def test_student_step_cuda_graph_replay(dev):
    """capturable=True: lr and the step count live in device memory, so one captured launch is replayed
    across steps and MultiStepLR milestones."""
    torch.manual_seed(9)
    cpu = [torch.randn(5000), torch.randn(77, 13)]
    student, teacher = Bag(cpu).to(dev), Bag(cpu).to(dev)
    opt = U.Adam(student.parameters(), lr=1e-2, capturable=True)
    tea = U.OldWeightEMA(teacher, student, alpha=0.9)
    opt.attach_teacher(tea)
    for p in student.parameters():
        p.grad = torch.zeros_like(p)
    s_cpu, t_cpu = [t.clone() for t in cpu], [t.clone() for t in cpu]
    m, v = [torch.zeros_like(t) for t in cpu], [torch.zeros_like(t) for t in cpu]
    g = torch.Generator().manual_seed(4)

    def feed():
        grads = [torch.randn(t.shape, generator=g) for t in cpu]
        for p, gr in zip(student.parameters(), grads):
            p.grad.copy_(gr)
        return grads

    grads = feed()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt.step(); tea.step()                      # eager warm-up step = update 1
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.no_grad():
        R.adam_step(s_cpu, grads, m, v, 1, 1e-2); R.ema_step(t_cpu, s_cpu, 0.9)
    graph = torch.cuda.CUDAGraph()
    grads = feed()
    with torch.cuda.graph(graph):
        opt.step(); tea.step()
    # capture does not execute: replay is update 2
    lr = 1e-2
    for it in range(2, 6):
        if it > 2:
            grads = feed()
        if it == 4:
            lr = 1e-3
            opt.param_groups[0]["lr"] = lr
            opt.sync_lr()
        graph.replay()
        with torch.no_grad():
            # the device lr is float32
            R.adam_step(s_cpu, grads, m, v, it, float(np.float32(lr))); R.ema_step(t_cpu, s_cpu, 0.9)
        assert_close_scaled(_cat(student.parameters()), _cat(s_cpu), RTOL, f"student replay {it}")
        assert_close_scaled(_cat(teacher.parameters()), _cat(t_cpu), RTOL, f"teacher replay {it}")
    assert opt.applied_steps() == 5


def test_grad_check_patterns(dev):
    torch.manual_seed(0)
    cpu = [torch.randn(4096 * 3 + 5), torch.randn(3), torch.randn(130, 7)]
    student = Bag(cpu).to(dev)
    opt = U.SGD(student.parameters(), lr=0.1)
    for p in student.parameters():
        p.grad = torch.randn_like(p)
    assert float(opt.check_grads()) == 0.0
    for ti, idx, val in ((0, 0, float("inf")), (0, 4096 * 3 + 4, float("nan")), (1, 2, float("-inf")), (2, 500, float("nan")),
                         (0, 5000, 3.0e38)):
        p = list(student.parameters())[ti]
        old = p.grad.view(-1)[idx].item()
        p.grad.view(-1)[idx] = val
        expect = 0.0 if np.isfinite(val) else 1.0
        for _ in range(2):   # the workspace words reset themselves
            assert float(opt.check_grads()) == expect, (ti, idx, val)
        p.grad.view(-1)[idx] = old
    assert float(opt.check_grads()) == 0.0


@pytest.mark.parametrize("algo", ["adam", "sgd"])
def test_state_dict_round_trip(dev, algo):
    """Checkpoint / resume (train_human.py:150-160,226-235 save and reload the optimizer state): an optimizer
    restored from state_dict() continues exactly like the uninterrupted one, including the device step count."""
    torch.manual_seed(13)
    cpu = [torch.randn(300), torch.randn(17, 9)]
    mk = (lambda ps: U.Adam(ps, lr=1e-2)) if algo == "adam" else (lambda ps: U.SGD(ps, lr=0.05, momentum=0.9, nesterov=True))
    g = torch.Generator().manual_seed(5)
    grads = [[torch.randn(t.shape, generator=g) for t in cpu] for _ in range(5)]

    def run(model, opt, its):
        for it in its:
            for p, gr in zip(model.parameters(), grads[it]):
                p.grad = gr.to(dev)
            opt.step()

    a = Bag(cpu).to(dev)
    opt_a = mk(a.parameters())
    run(a, opt_a, range(5))
    b = Bag(cpu).to(dev)
    opt_b = mk(b.parameters())
    run(b, opt_b, range(3))
    sd = opt_b.state_dict()
    assert all(float(st["step"]) == 3.0 for st in sd["state"].values())
    weights = [p.detach().clone() for p in b.parameters()]
    c = Bag([w.cpu() for w in weights]).to(dev)
    opt_c = mk(c.parameters())
    opt_c.load_state_dict(sd)
    assert opt_c.applied_steps() == 3
    run(c, opt_c, range(3, 5))
    assert opt_c.applied_steps() == 5
    for pa, pc in zip(a.parameters(), c.parameters()):
        assert torch.equal(pa.detach(), pc.detach())
    # torch's own optimizer accepts the same state (same keys)
    ref = (torch.optim.Adam if algo == "adam" else torch.optim.SGD)(c.parameters(), lr=1e-2, **({} if algo == "adam" else {"momentum": 0.9, "nesterov": True}))
    ref.load_state_dict(sd)


def test_sgd_resumes_a_torch_checkpoint_and_clones_per_parameter(dev):
    """torch.optim.SGD keeps no 'step' in its state and clones the gradient into the momentum buffer the first
    time EACH parameter is updated (`if buf is None`), also with dampening != 0.  (1) Resuming a torch SGD
    checkpoint must keep the loaded buffers (the old global `step <= 1` test overwrote them); (2) a parameter
    whose gradient first shows up at step 3 gets buf = grad then, the others keep dampening."""
    torch.manual_seed(3)
    cpu = [torch.randn(5000), torch.randn(33, 7), torch.randn(64)]
    kw = dict(lr=0.05, momentum=0.9, dampening=0.3, weight_decay=1e-3)
    g = torch.Generator().manual_seed(9)
    grads = [[torch.randn(t.shape, generator=g) for t in cpu] for _ in range(6)]

    def run(model, opt, its, late=None):
        for it in its:
            for i, (p, gr) in enumerate(zip(model.parameters(), grads[it])):
                p.grad = None if (late is not None and i == late and it < 3) else gr.to(dev)
            opt.step()

    # (1) torch writes the checkpoint after 2 steps, the drop-in resumes it
    ref = Bag(cpu).to(dev)
    ref_opt = torch.optim.SGD(ref.parameters(), **kw)
    run(ref, ref_opt, range(2))
    import copy
    sd = copy.deepcopy(ref_opt.state_dict())     # state_dict() hands out the LIVE buffers; load_state_dict keeps them
    assert all("step" not in st for st in sd["state"].values())
    new = Bag([p.detach().cpu() for p in ref.parameters()]).to(dev)
    new_opt = U.SGD(new.parameters(), **kw)
    new_opt.load_state_dict(sd)
    run(ref, ref_opt, range(2, 5))
    run(new, new_opt, range(2, 5))
    for a, b in zip(ref.parameters(), new.parameters()):
        assert_close_scaled(b.detach().cpu(), a.detach().cpu(), RTOL, "resumed SGD vs torch")
    # (2) parameter 1 has no gradient for three steps
    ref, new = Bag(cpu).to(dev), Bag(cpu).to(dev)
    ref_opt, new_opt = torch.optim.SGD(ref.parameters(), **kw), U.SGD(new.parameters(), **kw)
    run(ref, ref_opt, range(6), late=1)
    run(new, new_opt, range(6), late=1)
    for a, b in zip(ref.parameters(), new.parameters()):
        assert_close_scaled(b.detach().cpu(), a.detach().cpu(), RTOL, "SGD with a late gradient vs torch")
    for a, b in zip(ref.parameters(), new.parameters()):
        assert_close_scaled(new_opt.state[b]["momentum_buffer"].cpu(), ref_opt.state[a]["momentum_buffer"].cpu(), RTOL, "buffers")
    # (3) a skipped first step (found_inf) leaves the buffers "never written": the checkpoint carries none
    fresh = Bag(cpu).to(dev)
    fresh_opt = U.SGD(fresh.parameters(), **kw)
    for p, gr in zip(fresh.parameters(), grads[0]):
        p.grad = gr.to(dev)
    fresh_opt.grad_scale, fresh_opt.found_inf = torch.ones((), device=dev), torch.ones((), device=dev)
    fresh_opt.step()
    assert all("momentum_buffer" not in st for st in fresh_opt.state_dict()["state"].values())
    fresh_opt.found_inf = torch.zeros((), device=dev)
    fresh_opt.step()
    assert all("momentum_buffer" in st for st in fresh_opt.state_dict()["state"].values())
    ref1 = Bag(cpu).to(dev)
    ref1_opt = torch.optim.SGD(ref1.parameters(), **kw)
    run(ref1, ref1_opt, range(1))
    for a, b in zip(ref1.parameters(), fresh.parameters()):
        assert_close_scaled(b.detach().cpu(), a.detach().cpu(), RTOL, "first applied step after a skipped one")


def test_detach_teacher_restores_the_plain_ema(dev):
    """pretrain() steps the student optimizer without tea_optimizer.step() (train_human.py:243-300): after
    detach_teacher() the optimizer leaves the teacher alone and tea.step() is the plain EMA launch again."""
    cpu = [torch.randn(1000), torch.randn(40, 3)]
    s, t = Bag(cpu).to(dev), Bag(cpu).to(dev)
    opt = U.Adam(s.parameters(), lr=1e-2)
    tea = U.OldWeightEMA(t, s, alpha=0.9)
    opt.attach_teacher(tea)
    for p in s.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    tea.step()                                   # no-op: already folded in
    after_fused = _cat(t.parameters())
    assert not torch.equal(after_fused, _cat(cpu))
    opt.detach_teacher()
    opt.step()                                   # pretrain-style step: teacher untouched
    assert torch.equal(_cat(t.parameters()), after_fused)
    tea.step()                                   # and the EMA object works on its own again
    want = after_fused * 0.9 + _cat(s.parameters()) * (1 - 0.9)
    assert_close_scaled(_cat(t.parameters()), want, RTOL, "plain EMA after detach")

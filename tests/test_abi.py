"""C-ABI checks that need no GPU: the shared library builds/loads, exports every symbol that
include/udape.h declares, the ctypes prototype table covers exactly those symbols, argument
errors come back as negative codes with a message (no exception/abort across the ABI), and the
host-only chunk planner works.  No compute call is made here."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest
import torch

from uda_poseestimation_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "udape.h"


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"UDAPE_API\s+[\w\s\*]+?\b(udape_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ["udape_mean_std", "udape_adain_mix", "udape_decode", "udape_mask_select", "udape_pck_counts",
                 "udape_joints_mse_fwd", "udape_joints_mse_bwd", "udape_cons_fwd", "udape_cons_bwd",
                 "udape_gauss_target", "udape_labelmap", "udape_ema_plan", "udape_ema_multi",
                 "udape_version", "udape_last_error", "udape_build_info"]:
        assert must in syms, f"{must} missing from include/udape.h"


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"libudape_b200.so does not export {name}"
    assert sorted(_lib.PROTOTYPES) == declared_symbols(), "ctypes prototype table out of sync with udape.h"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.library_path()], capture_output=True, text=True)
    if out.returncode == 0:
        exported = {l.split()[-1] for l in out.stdout.splitlines() if " T " in l}
        assert set(declared_symbols()) <= exported


def test_version_and_build_info():
    lib = _lib.load()
    import re
    declared = int(re.search(r"#define UDAPE_VERSION (\d+)", (ROOT / "include" / "udape.h").read_text()).group(1))
    assert lib.udape_version() == declared >= 300
    info = lib.udape_build_info().decode()
    assert "sm_100a" in info and "udape-b200" in info


def test_header_is_plain_c(tmp_path):
    """the boundary must be consumable from C (no C++/torch types in the signatures)"""
    src = tmp_path / "t.c"
    src.write_text('#include "udape.h"\nint main(void){ udape_ema_chunk c; c.numel = 0; return (int)c.numel + (UDAPE_F32); }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o",
                        str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_sass_is_sm100a_only():
    r = subprocess.run(["cuobjdump", "--list-elf", _lib.library_path()], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", r.stdout))
    assert archs == {"sm_100a"}, archs


def test_argument_errors_are_negative_codes():
    lib = _lib.load()
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    # NULL tensor
    assert lib.udape_mean_std(None, _lib.F32, 4, 16, 1e-5, p, p, None) == -1
    assert "NULL" in _lib.last_error()
    # bad dtype
    assert lib.udape_mean_std(p, 9, 4, 16, 1e-5, p, p, None) == -2
    # bad shape
    assert lib.udape_adain_mix(p, p, _lib.F32, 0, 16, 16, 1e-5, 1.0, None, p, None) == -3
    # alpha outside [0,1] (Style_net.py:164)
    assert lib.udape_adain_mix(p, p, _lib.F32, 1, 16, 16, 1e-5, 1.5, None, p, None) == -5
    # misaligned pointer
    assert lib.udape_mean_std(p + 2, _lib.F32, 1, 4, 1e-5, p, p, None) == -4
    # kthvalue rank outside [1, n]
    assert lib.udape_mask_select(p, 10, 0, None, p, p, None) == -5
    assert lib.udape_mask_select(p, 10, 11, None, p, p, None) == -5
    assert lib.udape_decode(None, _lib.F32, 1, 4, 4, None, None, None, None, None, 0.0, None, 1.0, None, None) == -1
    assert lib.udape_labelmap(p, 1, 8, 8, 1.0, 7, 1, p, None, None) == -5
    with pytest.raises(ValueError):
        _lib.check(-3, "x")
    with pytest.raises(_lib.UdapeError):
        _lib.check(700, "x")


def test_ema_plan_host_only():
    lib = _lib.load()
    numel = [18, 4096, 4097, 10000, 0, 5]
    n_t = len(numel)
    base = 0x10000
    dst = (ctypes.c_void_p * n_t)(*[base + 0x100000 * i for i in range(n_t)])
    src = (ctypes.c_void_p * n_t)(*[base + 0x100000 * i + 0x80000 for i in range(n_t)])
    ne = (ctypes.c_int64 * n_t)(*numel)
    need = lib.udape_ema_plan(dst, src, ne, n_t, 4, 4096, None, 0)
    assert need == 1 + 1 + 2 + 3 + 0 + 1
    table = (_lib.EmaChunk * need)()
    assert lib.udape_ema_plan(dst, src, ne, n_t, 4, 4096, table, need) == need
    got = [(c.dst, c.src, c.numel) for c in table]
    assert got[0] == (base, base + 0x80000, 18)
    assert got[2] == (base + 0x200000, base + 0x280000, 4096)
    assert got[3] == (base + 0x200000 + 4096 * 4, base + 0x280000 + 4096 * 4, 1)
    assert sum(c[2] for c in got) == sum(numel)
    assert got[-1][2] == 5
    # errors: chunk size must keep 16-byte alignment; NULL tables
    assert lib.udape_ema_plan(dst, src, ne, n_t, 4, 1000, None, 0) < 0
    assert lib.udape_ema_plan(None, src, ne, n_t, 4, 4096, None, 0) < 0


def test_opt_plan_host_only_and_argument_errors():
    """udape_opt_plan is a host-only helper; udape_student_step / udape_grad_check / udape_decode_select /
    udape_mean_std_bwd validate their arguments before touching CUDA, so the error paths run without a GPU."""
    lib = _lib.load()
    numel = [18, 4096, 4097, 0, 5]
    n_t = len(numel)
    mk = lambda off, holes=(): (ctypes.c_void_p * n_t)(*[None if i in holes else 0x10000 + 0x100000 * i + off for i in range(n_t)])  # noqa: E731
    param, grad, m, v, ema = mk(0), mk(0x20000, holes=(1,)), mk(0x40000, holes=(1,)), mk(0x60000, holes=(1,)), mk(0x80000, holes=(4,))
    fresh = (ctypes.c_void_p * n_t)(*[None if i == 1 else 0x900000 + 4 * i for i in range(n_t)])
    ne = (ctypes.c_int64 * n_t)(*numel)
    need = lib.udape_opt_plan(param, grad, m, v, ema, fresh, ne, n_t, 4096, None, 0)
    assert need == 1 + 1 + 2 + 0 + 1
    table = (_lib.OptChunk * need)()
    assert lib.udape_opt_plan(param, grad, m, v, ema, fresh, ne, n_t, 4096, table, need) == need
    # the per-tensor "momentum buffer not written yet" word is the SAME address for every chunk of a tensor
    assert [c.fresh for c in table] == [0x900000, None, 0x900008, 0x900008, 0x900010]
    rows = [(c.param, c.grad, c.state1, c.state2, c.ema, c.numel) for c in table]
    assert rows[0] == (0x10000, 0x30000, 0x50000, 0x70000, 0x90000, 18)
    assert rows[1][1:4] == (None, None, None) and rows[1][4] == 0x110000 + 0x80000 and rows[1][5] == 4096   # no gradient: EMA only
    assert rows[3] == (0x210000 + 4096 * 4, 0x230000 + 4096 * 4, 0x250000 + 4096 * 4, 0x270000 + 4096 * 4, 0x290000 + 4096 * 4, 1)
    assert rows[4][4] is None and rows[4][5] == 5                                                           # no teacher: no EMA
    assert sum(r[5] for r in rows) == sum(numel)
    assert lib.udape_opt_plan(param, None, None, None, ema, None, ne, n_t, 4096, table, need) == need      # NULL tables allowed
    assert all(c.grad is None and c.state1 is None and c.fresh is None for c in table)
    assert lib.udape_opt_plan(None, grad, m, v, ema, None, ne, n_t, 4096, None, 0) == -1
    assert lib.udape_opt_plan(param, grad, m, v, ema, None, ne, n_t, 1000, None, 0) == -5
    # student step: argument validation
    h = _lib.OptHyper()
    h.lr, h.beta1, h.beta2, h.eps, h.step = 1e-3, 0.9, 0.999, 1e-8, 1
    fake = ctypes.c_void_p(0x1000)
    assert lib.udape_student_step(None, 0, _lib.OPT_ADAM, ctypes.byref(h), None, None, None, None, 0, None, 0, None, None) == 0   # nothing to do
    assert lib.udape_student_step(None, 4, _lib.OPT_ADAM, ctypes.byref(h), None, None, None, None, 0, None, 0, None, None) == -1
    assert lib.udape_student_step(fake, 4, 7, ctypes.byref(h), None, None, None, None, 0, None, 0, None, None) == -5
    h.step = 0
    assert lib.udape_student_step(fake, 4, _lib.OPT_ADAM, ctypes.byref(h), None, None, None, None, 0, None, 0, None, None) == -5
    h.step, h.beta1 = 1, 1.5
    assert lib.udape_student_step(fake, 4, _lib.OPT_ADAM, ctypes.byref(h), None, None, None, None, 0, None, 0, None, None) == -5
    h.beta1, h.beta2, h.nesterov = 0.0, 0.0, 1
    assert lib.udape_student_step(fake, 4, _lib.OPT_SGD, ctypes.byref(h), None, None, None, None, 0, None, 0, None, None) == -5   # nesterov without momentum
    assert "Nesterov" in _lib.last_error()
    h.beta1, h.nesterov = 0.9, 0
    assert lib.udape_student_step(fake, 4, _lib.OPT_SGD, ctypes.byref(h), None, None, None, None, 0, fake, 3, None, None) == -5   # flags without a ticket
    assert lib.udape_grad_check(fake, 4, None, None, None) == -1
    # decode_select / mean_std_bwd
    assert lib.udape_decode_select(fake, 0, 8, 4, 4, None, None, None, None, None, 0.0, None, 2.0, None, 3, None, None, None, fake, None) == -1
    assert lib.udape_decode_select(fake, 0, 8, 4, 4, None, None, None, fake, None, 0.0, None, 2.0, None, 9, None, None, None, fake, None) == -5
    assert lib.udape_mean_std_bwd(fake, fake, fake, None, None, 0, 0, 16, fake, None) == -3
    assert lib.udape_mean_std_bwd(fake, fake, fake, None, None, 3, 4, 16, fake, None) == -2


def test_fused_optimizer_host_side():
    """Constructor validation and torch.optim plumbing of the drop-in optimizers (no launch involved)."""
    import uda_poseestimation_b200 as U
    m = torch.nn.Linear(3, 2)
    with pytest.raises(ValueError):
        U.Adam(m.parameters(), lr=-1.0)
    with pytest.raises(ValueError):
        U.Adam(m.parameters(), betas=(1.0, 0.9))
    with pytest.raises(ValueError):
        U.SGD(m.parameters(), lr=0.1, nesterov=True)
    opt = U.SGD(m.parameters(), lr=0.1, momentum=0.9, nesterov=True)
    assert isinstance(opt, torch.optim.Optimizer) and opt._step_supports_amp_scaling
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [1], 0.1)      # train_human.py:143
    assert opt.param_groups[0]["lr"] == 0.1 and sched.get_last_lr() == [0.1]
    with pytest.raises(TypeError):
        opt.attach_teacher(object())
    with pytest.raises(RuntimeError, match="CUDA-only"):
        opt.step()
    assert issubclass(U.GradScaler, torch.amp.GradScaler)


def test_cpu_tensors_are_rejected_loudly():
    import uda_poseestimation_b200 as U
    x = torch.randn(2, 3, 8, 8)
    for call in (lambda: U.calc_mean_std(x), lambda: U.adain(x, x), lambda: U.JointsMSELoss()(x, x),
                 lambda: U.ConsLoss()(x, x), lambda: U.rectify(x, 2), lambda: U.get_max_preds_torch(x),
                 lambda: U.confidence_mask(x, 0.9), lambda: U.consistency_mask(x[:, :, 0, 0], 0.5)):
        with pytest.raises(RuntimeError, match="CUDA-only|no CUDA device"):
            call()


def test_product_package_does_not_import_the_oracle():
    """the oracle is test infrastructure: nothing under the package may reference it"""
    for f in (ROOT / "uda_poseestimation_b200").rglob("*.py"):
        text = f.read_text()
        assert "reference_port" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_dp_host_helpers_and_argument_errors():
    """The data-parallel entry points validate their arguments before touching CUDA; the slice geometry is a
    host-only helper: S = a multiple of 4096 elements, world slices cover the bucket, the last one may be short."""
    from uda_poseestimation_b200 import dp as DP
    lib = _lib.load()
    for n_total, world in [(52_991_568, 8), (52_991_568, 2), (4096, 8), (4100, 3), (64, 8), (8192 * 8, 8)]:
        s = lib.udape_dp_shard_elems(n_total, world)
        assert s % 4096 == 0 and s * world >= n_total and s * (world - 1) < max(n_total, 1) + 4096 * world
    assert lib.udape_dp_shard_elems(100, 0) == -5 and lib.udape_dp_shard_elems(100, 9) == -5
    offsets, total = DP.flat_layout([torch.zeros(3), torch.zeros(5, 2), torch.zeros(1), torch.zeros(8)])
    assert offsets == [0, 4, 16, 20] and total == 28                     # every tensor starts on a 16-byte boundary
    assert DP.arena_bytes(28) == DP.PAD_BYTES + 3 * 4 * 4096             # pad | grads | params | shadow, 4096-element granules
    peers = _lib.DpPeers()
    peers.rank, peers.world = 0, 2
    fake = ctypes.c_void_p(0x1000)
    h = _lib.OptHyper()
    h.lr, h.beta1, h.beta2, h.eps, h.step = 1e-3, 0.9, 0.999, 1e-8, 1
    # NULL pads / buffers
    assert lib.udape_dp_barrier(ctypes.byref(peers), 0, fake, 0, None) == -1
    for q in range(2):
        peers.pads[q] = 0x10000 + 0x2000 * q
    assert lib.udape_dp_barrier(ctypes.byref(peers), 9, fake, 0, None) == -5          # bad phase
    assert lib.udape_dp_reduce_step(ctypes.byref(peers), 4096, _lib.OPT_ADAM, ctypes.byref(h), None, None, fake, fake, fake, fake, fake, None) == -1  # no buffers
    for q in range(2):
        peers.grads[q], peers.params[q], peers.shadow[q] = 0x100000 + 0x10000 * q, 0x200000 + 0x10000 * q, 0x300000 + 0x10000 * q
    assert lib.udape_dp_reduce_step(ctypes.byref(peers), 4098, _lib.OPT_ADAM, ctypes.byref(h), None, None, fake, fake, fake, fake, fake, None) == -3  # n_total % 4
    assert lib.udape_dp_reduce_step(ctypes.byref(peers), 4096, _lib.OPT_ADAM, ctypes.byref(h), None, None, fake, None, None, fake, fake, None) == -1  # Adam state
    assert lib.udape_dp_gather_ema(ctypes.byref(peers), 4096, None, 0.9, 0.1, None, None, fake, fake, None) == -1                                   # step_dev
    assert lib.udape_dp_allreduce_counts(ctypes.byref(peers), fake, 65, fake, fake, 0, None) == -3                                                     # > 64 counts
    peers.world = 9
    assert lib.udape_dp_wait(ctypes.byref(peers), 1, fake, None, 0, None) == -5
    # multi-job entry points
    assert lib.udape_gauss_target_multi(None, 1, 8, 2.0, 256.0, 256.0, None) == -1
    jobs = (_lib.TargetJob * 1)()
    assert lib.udape_gauss_target_multi(jobs, 9, 8, 2.0, 256.0, 256.0, None) == -5
    assert lib.udape_gauss_target_multi(jobs, 1, 8, 2.0, 256.0, 256.0, None) == -1    # job with NULL pointers
    assert lib.udape_adain_mix_multi(None, 1, 0, 8, 16, 16, 1e-5, None) == -1
    assert lib.udape_table_feed(None, 1, 1, None, None, None) == -1

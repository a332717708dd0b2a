"""Pins the CPU oracle (oracle/reference_port.py) against the REAL reference functions, imported by file
path from the reference tree (oracle/ref_loader.py), on fresh random inputs — in addition to the committed
golden fixtures, which were produced by the same functions.  The tree exists only in the build container:
everywhere else (the GPU box) this module is skipped.  Equality is exact: the oracle calls the same
torch / numpy primitives in the same order.
"""
import os
import types

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import reference_port as R
from uda_poseestimation_b200 import synthetic as S

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present on this box")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_adain_and_statistics(seed):
    fn, sn = ref_loader.load("function"), ref_loader.load("style_net")
    g = torch.Generator().manual_seed(seed)
    c = torch.relu(torch.randn(3, 6, 17 + seed, 9, generator=g) + 0.2)
    s = torch.relu(torch.randn(3, 6, 8, 11 + seed, generator=g) * 2 + 0.5)
    for a, b in zip(R.calc_mean_std(c), fn.calc_mean_std(c)):
        assert torch.equal(a, b)
    assert torch.equal(R.adaptive_instance_normalization(c, s), fn.adaptive_instance_normalization(c, s))
    alpha = 0.1 + 0.4 * seed
    t = sn.adain(c, s)
    assert torch.equal(R.adain_mix(c, s, alpha), alpha * t + (1 - alpha) * c)      # Style_net.py:167-168


@pytest.mark.parametrize("seed", [3, 4])
def test_decode_accuracy_rectify(seed):
    kd, ut = ref_loader.load("keypoint_detection"), ref_loader.load("utils")
    hm = S.heatmaps(5, 7, seed=seed, peak=(0.2, 1.2))
    hm[0, 0] = 0.0
    hm[1, 2] = -1.0
    tgt = S.heatmaps(5, 7, seed=seed + 10)
    for x in (hm.numpy(), hm.half().numpy()):
        for a, b in zip(R.get_max_preds(x), kd.get_max_preds(x)):
            np.testing.assert_array_equal(a, b)
    for a, b in zip(R.get_max_preds_torch(hm), ut.get_max_preds_torch(hm)):
        assert torch.equal(a, b)
    ra, rb = R.accuracy(hm.numpy(), tgt.numpy()), kd.accuracy(hm.numpy(), tgt.numpy())
    np.testing.assert_array_equal(ra[0], rb[0])
    assert ra[1] == rb[1] and ra[2] == rb[2]
    np.testing.assert_array_equal(ra[3], rb[3])
    for sigma in (2, 1.0):
        assert torch.equal(R.rectify(hm.clone(), sigma), ut.rectify(hm.clone(), sigma))


@pytest.mark.parametrize("seed", [5, 6])
def test_losses(seed):
    lo = ref_loader.load("loss")
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(4, 5, 16, 16, generator=g).requires_grad_(True)
    o2 = o.detach().clone().requires_grad_(True)
    t = torch.rand(4, 5, 16, 16, generator=g)
    w = (torch.rand(4, 5, 1, generator=g) > 0.3).float()
    for red in ("mean", "none"):
        a, b = R.joints_mse_loss(o, t, w, reduction=red), lo.JointsMSELoss(reduction=red)(o2, t, w)
        assert torch.equal(a, b)
    a, b = R.joints_mse_loss(o, t, w), lo.JointsMSELoss()(o2, t, w)
    ga, gb = torch.autograd.grad(a, o)[0], torch.autograd.grad(b, o2)[0]
    assert torch.equal(ga, gb)
    m = torch.rand(4, 5, generator=g) > 0.5
    a, b = R.cons_loss(o, t, tea_mask=m), lo.ConsLoss()(o2, t, tea_mask=m)
    assert torch.equal(a, b)
    assert torch.equal(torch.autograd.grad(a, o)[0], torch.autograd.grad(b, o2)[0])


@pytest.mark.parametrize("sigma", [2, 1.0])
def test_target_heatmaps(sigma, capsys):
    du = ref_loader.load("dataset_util")
    joints, vis = S.keypoints(6, 9, seed=17)
    for i in range(6):
        a = R.generate_target(joints[i], vis[i], (64, 64), sigma, (256, 256))
        b = du.generate_target(joints[i], vis[i], (64, 64), sigma, (256, 256))
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
    pts = torch.tensor([[20.0, 30.0], [0.0, 0.0], [63.0, 63.0], [3.2, 60.9], [31.5, 31.5]])
    for p in pts:
        for kind in ("Gaussian", "Cauchy"):
            a = R.draw_labelmap_ori(torch.zeros(64, 64), p, sigma, type=kind)
            b = du.draw_labelmap_ori(torch.zeros(64, 64), p, sigma, type=kind)
            assert torch.equal(torch.as_tensor(a[0]), torch.as_tensor(b[0])) and a[1] == b[1]


def test_ema_and_student_step():
    ut = ref_loader.load("utils")

    def make(seed):
        torch.manual_seed(seed)
        return torch.nn.Sequential(torch.nn.Linear(9, 7), torch.nn.Linear(7, 3))

    teacher, student = make(1), make(2)
    tea = ut.OldWeightEMA(teacher, student, alpha=0.97)
    opt = torch.optim.Adam(student.parameters(), lr=2e-3)
    s_cpu = [p.detach().clone() for p in student.parameters()]
    t_cpu = [p.detach().clone() for p in teacher.parameters()]
    m = [torch.zeros_like(p) for p in s_cpu]
    v = [torch.zeros_like(p) for p in s_cpu]
    g = torch.Generator().manual_seed(3)
    step = 0
    for _ in range(4):
        grads = [torch.randn(p.shape, generator=g) for p in s_cpu]
        for p, gr in zip(student.parameters(), grads):
            p.grad = gr.clone()
        opt.step()        # train_human.py:437 without a scaler
        tea.step()        # :438
        _, step = R.student_teacher_step("adam", s_cpu, grads, m, v, t_cpu, step, None, 0.97,
                                         lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
        for a, b in zip(s_cpu + t_cpu, list(student.parameters()) + list(teacher.parameters())):
            assert torch.equal(a, b.detach())


def test_style_loss():
    net = ref_loader.load("adain_net")
    fake_self = types.SimpleNamespace(mse_loss=torch.nn.MSELoss())
    g = torch.Generator().manual_seed(8)
    x = torch.relu(torch.randn(2, 4, 24, 24, generator=g)).requires_grad_(True)
    x2 = x.detach().clone().requires_grad_(True)
    t = torch.relu(torch.randn(2, 4, 24, 24, generator=g) * 1.3)
    a, b = R.calc_style_loss(x, t), net.Net.calc_style_loss(fake_self, x2, t)
    assert torch.equal(a, b)
    assert torch.equal(torch.autograd.grad(a, x)[0], torch.autograd.grad(b, x2)[0])


def test_style_transfer_forward_only():
    """The part of Style_net.Net.forward the trainers keep (element [2]) + the clamp of train_human.py:276,
    against the real class with the reference module's own (randomly initialised) VGG-19 / decoder."""
    sn = ref_loader.load("style_net")
    torch.manual_seed(0)
    net = sn.Net(sn.vgg, sn.decoder).eval()
    g = torch.Generator().manual_seed(1)
    content, style = torch.randn(2, 3, 32, 32, generator=g), torch.randn(2, 3, 32, 32, generator=g)
    lo, hi = torch.tensor([-2.1179, -2.0357, -1.8044]), torch.tensor([2.2489, 2.4285, 2.64])
    with torch.no_grad():
        g_t = net(content, style, 0.6)[2]
        ref = torch.maximum(torch.minimum(g_t.permute(0, 2, 3, 1), hi), lo).permute(0, 3, 1, 2)
    assert torch.equal(R.style_transfer(sn.vgg, sn.decoder, content, style, 0.6), g_t)
    assert torch.equal(R.style_transfer(sn.vgg, sn.decoder, content, style, 0.6, lo, hi), ref)


def test_bytecode_of_the_reference_loads_without_the_tree():
    """oracle/build_ref.py compiles the hot-path files into oracle/_ref/*.code (.pyc format, outputs only); a process that cannot
    see /root/reference — the GPU box — imports the reference's functions from there and gets the reference's results."""
    import subprocess
    import sys
    from pathlib import Path

    from oracle import build_ref

    if not build_ref.build(verbose=False):
        pytest.skip("no reference tree and no prebuilt bytecode")
    root = Path(__file__).resolve().parents[1]
    code = (
        "import torch, numpy as np\n"
        "from oracle import ref_loader as L, reference_live as RL, reference_port as R\n"
        "from uda_poseestimation_b200 import synthetic as S\n"
        "assert L.kind() == 'bytecode', L.kind()\n"
        "assert L.load('utils').rectify.__code__.co_filename.startswith('<reference>/')\n"
        "hm = S.heatmaps(3, 4, seed=3, peak=(0.3, 1.2))\n"
        "assert torch.equal(RL.rectify(hm.clone(), 2), R.rectify(hm.clone(), 2))\n"
        "c, s = S.vgg_features(2, seed=1, channels=8)\n"
        "assert torch.equal(RL.adain_mix(c, s, 0.3), R.adain_mix(c, s, 0.3))\n"
        "a, b = RL.accuracy(hm.numpy(), S.heatmaps(3, 4, seed=4).numpy()), R.accuracy(hm.numpy(), S.heatmaps(3, 4, seed=4).numpy())\n"
        "assert a[1] == b[1] and a[2] == b[2] and np.array_equal(a[0], b[0])\n"
        "j, v = S.keypoints(2, 5, seed=2)\n"
        "assert np.array_equal(RL.generate_target(j[0], v[0], (64, 64), 2, (256, 256))[0], R.generate_target(j[0], v[0], (64, 64), 2, (256, 256))[0])\n"
        "print('ok')\n"
    )
    env = dict(os.environ, UDAPE_REFERENCE_ROOT="/nonexistent-reference-root")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_reference_bytecode_manifest(tmp_path, monkeypatch):
    """oracle/build_ref.py: a second build with unchanged sources recompiles nothing; bytecode of another interpreter
    version (a different magic number in the manifest) is not used."""
    import json

    from oracle import build_ref

    if not ref_loader.source_available():
        pytest.skip("needs the reference tree to compile from")
    monkeypatch.setattr(build_ref, "OUT", tmp_path)
    monkeypatch.setattr(ref_loader, "BYTECODE_ROOT", tmp_path)
    assert build_ref.build(verbose=False) and ref_loader.bytecode_available()
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == sorted(["MANIFEST.json"] + [f"{n}.code" for n in ref_loader._FILES])
    for p in tmp_path.iterdir():       # outputs only: marshalled code objects, no source text of the reference
        assert b"def calc_mean_std(feat" not in p.read_bytes()
    stamps = {p.name: p.stat().st_mtime_ns for p in tmp_path.glob("*.code")}
    assert build_ref.build(verbose=False)
    assert stamps == {p.name: p.stat().st_mtime_ns for p in tmp_path.glob("*.code")}
    manifest = json.loads((tmp_path / "MANIFEST.json").read_text())
    assert set(manifest["files"]) == set(ref_loader._FILES) and all(len(v["sha256"]) == 64 for v in manifest["files"].values())
    manifest["magic"] = "00000000"
    (tmp_path / "MANIFEST.json").write_text(json.dumps(manifest))
    assert not ref_loader.bytecode_available()
    assert build_ref.build(verbose=False) and ref_loader.bytecode_available()      # rebuilt for this interpreter

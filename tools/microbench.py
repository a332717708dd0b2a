#!/usr/bin/env python
"""Per-kernel HBM-bandwidth microbenchmark (BASELINE.json configs[4] and the per-config sizes).

Each kernel is launched directly through the C-ABI on pre-allocated buffers, `--iters` times,
with the whole L2 flushed (a 512 MB memset) before every timed launch and a CUDA-event pair
around the launch.  Reported: median launch time, achieved GB/s = algorithmic bytes / time, and
the fraction of the measured copy peak (MEASURED_PEAKS.json) and of the 8 TB/s spec.

    python tools/microbench.py [--iters 20] [--out gpurun_out/microbench.json] [--only adain,ema]

Event pairs add ~2-3 us, so kernels that move < ~30 MB are launch/latency-bound here; the ncu
launch list under profiles/ gives their pure device time.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from uda_poseestimation_b200 import _lib  # noqa: E402
from uda_poseestimation_b200 import synthetic as S  # noqa: E402
from uda_poseestimation_b200.ema import MultiTensorPlan  # noqa: E402

DT = {"f32": (torch.float32, _lib.F32), "f16": (torch.float16, _lib.F16), "bf16": (torch.bfloat16, _lib.BF16)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/microbench.json")
    ap.add_argument("--only", default="")
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    peaks = ROOT / "MEASURED_PEAKS.json"
    peak = float(json.loads(peaks.read_text())["hbm_gbs"]) if peaks.exists() else 6650.0
    rows = []

    def bench(name, shape, nbytes, fn, group):
        if only and group not in only:
            return
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
        for a, b in evs:
            if not args.no_flush:
                flush_buf.zero_()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ts = np.array([a.elapsed_time(b) for a, b in evs])
        med = float(np.median(ts))
        gbs = nbytes / (med * 1e-3) / 1e9
        rows.append(dict(kernel=name, shape=shape, mbytes=nbytes / 1e6, us=med * 1e3, us_min=float(ts.min()) * 1e3,
                         gbs=gbs, frac_measured=gbs / peak, frac_spec=gbs / 8000.0))
        print(f"{name:<28}{shape:<26}{nbytes / 1e6:9.1f} MB {med * 1e3:9.1f} us {gbs:8.0f} GB/s "
              f"{100 * gbs / peak:6.1f}% meas {100 * gbs / 8000:6.1f}% spec", flush=True)

    def chk(s):
        _lib.check(s, "microbench")

    # ---- AdaIN statistics and fused AdaIN+mix: N x 512 x 32 x 32 --------------------------------------
    for n in (8, 32, 64, 128):
        for dn in (("f32", "bf16") if n == 32 else ("f32",)):
            td, code = DT[dn]
            c = torch.relu(torch.randn(n, 512, 32, 32, device=dev)).to(td)
            s_ = torch.relu(torch.randn(n, 512, 32, 32, device=dev) * 2).to(td)
            out = torch.empty_like(c)
            mean = torch.empty(n * 512, dtype=td, device=dev)
            std = torch.empty_like(mean)
            e = c.element_size()
            bench("mean_std", f"{n}x512x32x32 {dn}", c.numel() * e + 2 * n * 512 * e,
                  lambda: chk(lib.udape_mean_std(c.data_ptr(), code, n * 512, 1024, 1e-5, mean.data_ptr(), std.data_ptr(), st())),
                  "adain")
            bench("adain_mix", f"{n}x512x32x32 {dn}", 3 * c.numel() * e,
                  lambda: chk(lib.udape_adain_mix(c.data_ptr(), s_.data_ptr(), code, n * 512, 1024, 1024, 1e-5, 0.5, None,
                                                  out.data_ptr(), st())), "adain")
            del c, s_, out

    # ---- heatmap kernels ---------------------------------------------------------------------------------
    for cfg in ("C1", "C2", "C4", "C5"):
        b, k, sigma = S.CONFIGS[cfg]["batch"], S.CONFIGS[cfg]["joints"], S.CONFIGS[cfg]["sigma"]
        planes, hw = b * k, 4096
        tea = S.heatmaps(b, k, seed=1, peak=(0.3, 1.2)).to(dev)
        stu16 = S.heatmaps(b, k, seed=2).to(dev).half()
        label = S.heatmaps(b, k, seed=3, noise=0.0).to(dev)
        weight = torch.ones(planes, device=dev)
        rect = torch.empty_like(tea)
        preds = torch.empty(planes, 2, device=dev)
        maxv = torch.empty(planes, device=dev)
        pos = torch.empty(planes, 2, dtype=torch.int64, device=dev)
        conf = torch.empty(planes, dtype=torch.uint8, device=dev)
        shape = f"{cfg} {b}x{k}x64x64"
        bench("decode f32", shape, planes * hw * 4 + 16 * planes,
              lambda: chk(lib.udape_decode(tea.data_ptr(), _lib.F32, planes, 64, 64, None, preds.data_ptr(), maxv.data_ptr(),
                                           None, None, 0.9, None, 2.0, None, st())), "decode")
        bench("decode f16", shape, planes * hw * 2 + 16 * planes,
              lambda: chk(lib.udape_decode(stu16.data_ptr(), _lib.F16, planes, 64, 64, None, preds.data_ptr(), None,
                                           maxv.data_ptr(), None, 0.9, None, 2.0, None, st())), "decode")
        bench("decode+conf+rectify f32", shape, 2 * planes * hw * 4 + 32 * planes,
              lambda: chk(lib.udape_decode(tea.data_ptr(), _lib.F32, planes, 64, 64, None, None, None, maxv.data_ptr(),
                                           pos.data_ptr(), 0.9, conf.data_ptr(), float(sigma), rect.data_ptr(), st())), "decode")
        thresh = torch.empty(1, device=dev)
        tm = torch.empty(planes, dtype=torch.uint8, device=dev)
        bench("mask_select", shape, 9 * planes,
              lambda: chk(lib.udape_mask_select(maxv.data_ptr(), planes, planes // 2, None, thresh.data_ptr(), tm.data_ptr(), st())),
              "decode")
        hits = torch.empty(2, k, dtype=torch.int32, device=dev)
        bench("pck f16/f32", shape, planes * hw * 6 + 8 * planes,
              lambda: chk(lib.udape_pck_counts(stu16.data_ptr(), _lib.F16, label.data_ptr(), _lib.F32, b, k, 64, 64, 0.5,
                                               preds.data_ptr(), None, hits[0].data_ptr(), hits[1].data_ptr(), st())), "pck")
        scratch = torch.empty(planes + 4, device=dev)
        base = scratch.data_ptr()
        g1 = torch.full((1,), 65536.0, device=dev)
        grad16 = torch.empty_like(stu16)
        bench("joints_mse_fwd f16/f32", shape, planes * hw * 6 + 4 * planes,
              lambda: chk(lib.udape_joints_mse_fwd(stu16.data_ptr(), _lib.F16, label.data_ptr(), _lib.F32, weight.data_ptr(),
                                                   _lib.F32, planes, hw, base, base + 4 * planes, base + 4 * planes + 4, st())),
              "loss")
        bench("joints_mse_bwd f16/f32", shape, planes * hw * 8,
              lambda: chk(lib.udape_joints_mse_bwd(stu16.data_ptr(), _lib.F16, label.data_ptr(), _lib.F32, weight.data_ptr(),
                                                   _lib.F32, planes, hw, g1.data_ptr(), 0, grad16.data_ptr(), st())), "loss")
        bench("cons_fwd f16/f32", shape, planes * hw * 6 + 4 * planes,
              lambda: chk(lib.udape_cons_fwd(stu16.data_ptr(), _lib.F16, rect.data_ptr(), _lib.F32, tm.data_ptr(), _lib.U8,
                                             None, b, k, hw, base, base + 4 * planes + 8, base + 4 * planes,
                                             base + 4 * planes + 4, st())), "loss")
        bench("cons_bwd f16/f32", shape, planes * hw * 8,
              lambda: chk(lib.udape_cons_bwd(stu16.data_ptr(), _lib.F16, rect.data_ptr(), _lib.F32, tm.data_ptr(), _lib.U8,
                                             None, b, k, hw, g1.data_ptr(), None, grad16.data_ptr(), st())), "loss")
        joints, vis = S.keypoints(b, k, seed=4)
        jd = torch.from_numpy(joints).to(dev).reshape(planes, 2).contiguous()
        vd = torch.from_numpy(vis).to(dev).reshape(planes).contiguous()
        wout = torch.empty(planes, device=dev)
        bench("gauss_target", shape, planes * hw * 4 + 24 * planes,
              lambda: chk(lib.udape_gauss_target(jd.data_ptr(), vd.data_ptr(), planes, 64, 64, float(sigma), 256.0, 256.0,
                                                 rect.data_ptr(), wout.data_ptr(), st())), "target")
        pts = (jd / 4).to(torch.int32).contiguous()
        visout = torch.empty(planes, dtype=torch.int32, device=dev)
        bench("labelmap", shape, planes * hw * 4 + 12 * planes,
              lambda: chk(lib.udape_labelmap(pts.data_ptr(), planes, 64, 64, float(sigma), 0, 1, rect.data_ptr(),
                                             visout.data_ptr(), st())), "target")

    # ---- EMA over the PoseResNet-101 parameter census -----------------------------------------------------
    shapes = S.pose_resnet_param_shapes(21)
    student = S.parameter_list(shapes, 1, device=dev)
    teacher = [t.clone() for t in student]
    plan = MultiTensorPlan(teacher, student)
    n_params = sum(t.numel() for t in student)
    bench("ema_multi f32", f"PoseResNet-101 {n_params}", 3 * n_params * 4, lambda: plan.run(0.999, 0.001, 0), "ema")
    plan_c = MultiTensorPlan(teacher, student, as_bytes=True)
    bench("ema_multi copy", f"PoseResNet-101 {n_params}", 2 * n_params * 4, lambda: plan_c.run(0.0, 1.0, 1), "ema")
    # reference point: torch's own copy of the same bytes (what MEASURED_PEAKS measures)
    big_a = torch.empty(n_params, device=dev)
    big_b = torch.empty(n_params, device=dev)
    bench("torch copy_ (reference pt)", f"{n_params} f32", 2 * n_params * 4, lambda: big_b.copy_(big_a), "ema")

    out = Path(args.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(dict(peak_gbs=peak, device=torch.cuda.get_device_name(0), iters=args.iters,
                                   l2_flush=not args.no_flush, rows=rows), indent=1))
    print(f"wrote {out}")


if __name__ == "__main__":
    main()

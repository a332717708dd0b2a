#!/usr/bin/env python
"""Per-kernel HBM-bandwidth microbenchmark (BASELINE.json configs[4] and the per-config sizes).

Each kernel is launched directly through the C-ABI on pre-allocated buffers and timed two ways:

  sustained  `--iters` x R launches captured into ONE CUDA graph and replayed between an event pair, rotating over
             R independent buffer sets whose total footprint exceeds 2x the 126 MB L2 (so every launch
             streams from HBM); launch latency is hidden behind the previous kernel, as it is inside
             the step's CUDA graph.  This is the roofline figure: GB/s = algorithmic bytes / avg time.
  isolated   one launch between an event pair after a whole-L2 flush (512 MB memset): adds the
             ~2-4 us launch + event overhead, i.e. the latency a lone call sees.

Fractions are of the measured copy peak (MEASURED_PEAKS.json) and of the 8 TB/s spec.

    python tools/microbench.py [--iters 20] [--out gpurun_out/microbench.json] [--only adain,ema]
"""
from __future__ import annotations

import argparse
import ctypes
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
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-sustained", action="store_true", help="isolated launches only (use under ncu)")
    ap.add_argument("--adain-n", default="8,32,64,128", help="feature batch sizes for the AdaIN kernels")
    ap.add_argument("--configs", default="C1,C2,C4,C5", help="heatmap configs to run")
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
    L2_BYTES = 126 << 20

    def bench(name, shape, nbytes, make, group, footprint=None):
        """make(r) -> closure launching the kernel on buffer set r (allocates its own tensors)."""
        if only and group not in only:
            return
        fp = footprint or nbytes
        sets = max(1, min(40, -(-(2 * L2_BYTES) // fp)))
        fns = [make(r) for r in range(sets)]
        for _ in range(args.warmup):
            for fn in fns:
                fn()
        torch.cuda.synchronize()
        # isolated: flush + one launch per event pair
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
        for i, (a, b) in enumerate(evs):
            if not args.no_flush:
                flush_buf.zero_()
            a.record()
            fns[i % sets]()
            b.record()
        torch.cuda.synchronize()
        ts = np.array([a.elapsed_time(b) for a, b in evs])
        iso = float(np.median(ts))
        # sustained: back-to-back launches over the rotating buffer sets, replayed as ONE CUDA graph so
        # that the host (Python + ctypes, ~8 us per call) is not what is being measured
        reps = max(1, args.iters)
        if args.no_sustained:
            rows.append(dict(kernel=name, shape=shape, mbytes=nbytes / 1e6, us_isolated=iso * 1e3))
            print(f"{name:<28}{shape:<26}{nbytes / 1e6:9.1f} MB isolated {iso * 1e3:7.1f} us", flush=True)
            return
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for _ in range(reps):
                    for fn in fns:
                        fn()
        torch.cuda.current_stream().wait_stream(side)
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for _ in range(3):
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            t = a.elapsed_time(b) / (reps * sets)
            best = t if best is None else min(best, t)
        del graph
        gbs = nbytes / (best * 1e-3) / 1e9
        rows.append(dict(kernel=name, shape=shape, mbytes=nbytes / 1e6, us=best * 1e3, us_isolated=iso * 1e3,
                         buffer_sets=sets, gbs=gbs, frac_measured=gbs / peak, frac_spec=gbs / 8000.0,
                         gbs_isolated=nbytes / (iso * 1e-3) / 1e9))
        print(f"{name:<28}{shape:<26}{nbytes / 1e6:9.1f} MB {best * 1e3:8.1f} us {gbs:7.0f} GB/s "
              f"{100 * gbs / peak:6.1f}% meas {100 * gbs / 8000:5.1f}% spec | isolated {iso * 1e3:7.1f} us x{sets}", flush=True)

    def chk(s):
        _lib.check(s, "microbench")

    # ---- AdaIN statistics and fused AdaIN+mix: N x 512 x 32 x 32 --------------------------------------
    for n in [int(x) for x in args.adain_n.split(",") if x]:
        for dn in (("f32", "bf16") if n == 32 else ("f32",)):
            td, code = DT[dn]
            cache = {}

            def feat(r, n=n, td=td, cache=cache):
                if r not in cache:
                    c = torch.relu(torch.randn(n, 512, 32, 32, device=dev)).to(td)
                    s_ = torch.relu(torch.randn(n, 512, 32, 32, device=dev) * 2).to(td)
                    cache[r] = (c, s_, torch.empty_like(c), torch.empty(n * 512, dtype=td, device=dev),
                                torch.empty(n * 512, dtype=td, device=dev))
                return cache[r]

            e = 4 if dn == "f32" else 2
            numel = n * 512 * 1024

            def mk_ms(r, n=n, code=code, feat=feat):
                c, _, _, mean, std = feat(r)
                return lambda: chk(lib.udape_mean_std(c.data_ptr(), code, n * 512, 1024, 1e-5, mean.data_ptr(), std.data_ptr(), st()))

            def mk_ad(r, n=n, code=code, feat=feat):
                c, s_, out, _, _ = feat(r)
                return lambda: chk(lib.udape_adain_mix(c.data_ptr(), s_.data_ptr(), code, n * 512, 1024, 1024, 1e-5, 0.5, None,
                                                       out.data_ptr(), st()))

            bench("mean_std", f"{n}x512x32x32 {dn}", numel * e + 2 * n * 512 * e, mk_ms, "adain")
            bench("adain_mix", f"{n}x512x32x32 {dn}", 3 * numel * e, mk_ad, "adain")

            def mk_ad2(r, n=n, code=code, feat=feat):   # both directions of a step (s2t + t2s) in one launch
                c, s_, out, _, _ = feat(r)
                c2, s2, out2, _, _ = feat(r + 1000)
                jobs = (_lib.AdainJob * 2)()
                for i, (a_, b_, o_) in enumerate(((c, s_, out), (c2, s2, out2))):
                    jobs[i].content, jobs[i].style, jobs[i].out, jobs[i].alpha_dev, jobs[i].alpha = a_.data_ptr(), b_.data_ptr(), o_.data_ptr(), None, 0.5
                return lambda: chk(lib.udape_adain_mix_multi(jobs, 2, code, n * 512, 1024, 1024, 1e-5, st()))

            if n <= 64:
                bench("adain_mix x2 (1 launch)", f"{n}x512x32x32 {dn}", 6 * numel * e, mk_ad2, "adain")
            cache.clear()

    # ---- heatmap kernels ---------------------------------------------------------------------------------
    for cfg in [c for c in args.configs.split(",") if c]:
        b, k, sigma = S.CONFIGS[cfg]["batch"], S.CONFIGS[cfg]["joints"], S.CONFIGS[cfg]["sigma"]
        planes, hw = b * k, 4096
        base_tea = S.heatmaps(b, k, seed=1, peak=(0.3, 1.2)).to(dev)
        base_stu = S.heatmaps(b, k, seed=2).to(dev).half()
        base_label = S.heatmaps(b, k, seed=3, noise=0.0).to(dev)
        joints, vis = S.keypoints(b, k, seed=4)
        jd = torch.from_numpy(joints).to(dev).reshape(planes, 2).contiguous()
        vd = torch.from_numpy(vis).to(dev).reshape(planes).contiguous()
        pts = (jd / 4).to(torch.int32).contiguous()
        g1 = torch.full((1,), 65536.0, device=dev)
        weight = torch.ones(planes, device=dev)
        bundles = {}

        def B(r, bundles=bundles, planes=planes, k=k):
            """buffer set r of this config (inputs are clones so that every set has its own addresses)"""
            if r not in bundles:
                d = dict(tea=base_tea.clone(), stu16=base_stu.clone(), stu16b=base_stu.clone(), label=base_label.clone(),
                         rect=torch.empty_like(base_tea), grad16=torch.empty_like(base_stu), grad16b=torch.empty_like(base_stu),
                         preds=torch.empty(planes, 2, device=dev), maxv=torch.rand(planes, device=dev),
                         pos=torch.empty(planes, 2, dtype=torch.int64, device=dev),
                         conf=torch.empty(planes, dtype=torch.uint8, device=dev), thresh=torch.empty(1, device=dev),
                         tm=torch.ones(planes, dtype=torch.uint8, device=dev),
                         hits=torch.empty(2, k, dtype=torch.int32, device=dev),
                         scratch=torch.empty(2 * planes + 8, device=dev), wout=torch.empty(planes, device=dev),
                         visout=torch.empty(planes, dtype=torch.int32, device=dev))
                bundles[r] = d
            return bundles[r]

        shape = f"{cfg} {b}x{k}x64x64"
        hm32, hm16 = planes * hw * 4, planes * hw * 2

        def mk_dec32(r):
            d = B(r)
            return lambda: chk(lib.udape_decode(d["tea"].data_ptr(), _lib.F32, planes, 64, 64, None, d["preds"].data_ptr(),
                                                d["maxv"].data_ptr(), None, None, 0.9, None, 2.0, None, st()))

        def mk_dec16(r):
            d = B(r)
            return lambda: chk(lib.udape_decode(d["stu16"].data_ptr(), _lib.F16, planes, 64, 64, None, d["preds"].data_ptr(), None,
                                                d["maxv"].data_ptr(), None, 0.9, None, 2.0, None, st()))

        def mk_decrect(r):
            d = B(r)
            return lambda: chk(lib.udape_decode(d["tea"].data_ptr(), _lib.F32, planes, 64, 64, None, None, None, d["maxv"].data_ptr(),
                                                d["pos"].data_ptr(), 0.9, d["conf"].data_ptr(), float(sigma), d["rect"].data_ptr(), st()))

        def mk_decsel(r):   # the step's teacher call: decode + conf_table + k-th value mask in one launch, no map written
            d = B(r)
            tk = _lib.ticket(dev)
            return lambda: chk(lib.udape_decode_select(d["tea"].data_ptr(), _lib.F32, planes, 64, 64, None, d["preds"].data_ptr(), None,
                                                       d["maxv"].data_ptr(), d["pos"].data_ptr(), 0.9, d["conf"].data_ptr(), float(sigma), None,
                                                       planes // 2, None, d["thresh"].data_ptr(), d["tm"].data_ptr(), tk, st()))

        def mk_sel(r):
            d = B(r)
            return lambda: chk(lib.udape_mask_select(d["maxv"].data_ptr(), planes, planes // 2, None, d["thresh"].data_ptr(),
                                                     d["tm"].data_ptr(), st()))

        def mk_pck(r):
            d = B(r)
            tk = _lib.ticket(dev)
            return lambda: chk(lib.udape_pck_counts(d["stu16"].data_ptr(), _lib.F16, d["label"].data_ptr(), _lib.F32, b, k, 64, 64, 0.5,
                                                    d["preds"].data_ptr(), None, d["hits"][0].data_ptr(), d["hits"][1].data_ptr(), d["conf"].data_ptr(), tk, st()))

        def mk_msef(r):
            d = B(r)
            base = d["scratch"].data_ptr()
            tk = _lib.ticket(dev)
            return lambda: chk(lib.udape_joints_mse_fwd(d["stu16"].data_ptr(), _lib.F16, d["label"].data_ptr(), _lib.F32, weight.data_ptr(),
                                                        _lib.F32, planes, hw, base, base + 4 * planes, tk, st()))

        def mk_mseb(r):
            d = B(r)
            return lambda: chk(lib.udape_joints_mse_bwd(d["stu16"].data_ptr(), _lib.F16, d["label"].data_ptr(), _lib.F32, weight.data_ptr(),
                                                        _lib.F32, planes, hw, g1.data_ptr(), 0, d["grad16"].data_ptr(), st()))

        def mk_consf(r):
            d = B(r)
            base = d["scratch"].data_ptr()
            tk = _lib.ticket(dev)
            return lambda: chk(lib.udape_cons_fwd(d["stu16"].data_ptr(), _lib.F16, d["tea"].data_ptr(), _lib.F32, d["tm"].data_ptr(), _lib.U8,
                                                  None, b, k, hw, base, base + 4 * planes + 8, base + 4 * planes, tk, st()))

        def mk_consb(r):
            d = B(r)
            return lambda: chk(lib.udape_cons_bwd(d["stu16"].data_ptr(), _lib.F16, d["tea"].data_ptr(), _lib.F32, d["tm"].data_ptr(), _lib.U8,
                                                  None, b, k, hw, g1.data_ptr(), None, d["grad16"].data_ptr(), st()))

        def mk_step(r, analytic):
            d = B(r)
            base = d["scratch"].data_ptr()
            tk = _lib.ticket(dev)
            # supervised pair (stu16 vs label) + consistency pair (stu16 vs teacher map | analytic from preds)
            return lambda: chk(lib.udape_loss_step(
                d["stu16"].data_ptr(), d["label"].data_ptr(), weight.data_ptr(), _lib.F32, planes, d["stu16b"].data_ptr(),
                None if analytic else d["tea"].data_ptr(), d["preds"].data_ptr() if analytic else None, float(sigma),
                d["tm"].data_ptr(), _lib.U8, b, k, 64, 64, _lib.F16, _lib.F32, 1.0, 65536.0, None, base,
                base + 8 * planes, tk, d["grad16"].data_ptr(), d["grad16b"].data_ptr(), st()))

        def mk_gt(r):
            d = B(r)
            return lambda: chk(lib.udape_gauss_target(jd.data_ptr(), vd.data_ptr(), planes, 64, 64, float(sigma), 256.0, 256.0,
                                                      d["rect"].data_ptr(), d["wout"].data_ptr(), st()))

        def mk_lm(r):
            d = B(r)
            return lambda: chk(lib.udape_labelmap(pts.data_ptr(), planes, 64, 64, float(sigma), 0, 1, d["rect"].data_ptr(),
                                                  d["visout"].data_ptr(), st()))

        bench("decode f32", shape, hm32 + 16 * planes, mk_dec32, "decode")
        bench("decode f16", shape, hm16 + 16 * planes, mk_dec16, "decode")
        bench("decode+conf+rectify f32", shape, 2 * hm32 + 32 * planes, mk_decrect, "decode")
        bench("decode+conf+kth-select f32", shape, hm32 + 41 * planes, mk_decsel, "decode")
        # valid arg-max coordinates for the analytic loss step
        mk_dec32(0)()
        for r_ in list(bundles):
            bundles[r_]["preds"].copy_(bundles[0]["preds"])
        bench("mask_select", shape, 9 * planes, mk_sel, "decode", footprint=1 << 30)
        bench("pck f16/f32", shape, hm16 + hm32 + 8 * planes, mk_pck, "pck")
        bench("joints_mse_fwd f16/f32", shape, hm16 + hm32 + 4 * planes, mk_msef, "loss")
        bench("joints_mse_bwd f16/f32", shape, 2 * hm16 + hm32, mk_mseb, "loss")
        bench("cons_fwd f16/f32", shape, hm16 + hm32 + 4 * planes, mk_consf, "loss")
        bench("cons_bwd f16/f32", shape, 2 * hm16 + hm32, mk_consb, "loss")
        for r_ in list(bundles):
            bundles[r_]["preds"].copy_(bundles[0]["preds"])
        bench("loss_step (mse+cons, maps)", shape, 2 * (2 * hm16 + hm32), lambda r: mk_step(r, False), "loss")
        bench("loss_step (analytic teacher)", shape, 2 * (2 * hm16) + hm32, lambda r: mk_step(r, True), "loss")
        bench("gauss_target", shape, hm32 + 24 * planes, mk_gt, "target")
        bench("labelmap", shape, hm32 + 12 * planes, mk_lm, "target")

        # every target set of a loader in one launch: three 64x64 + two 8x8 sets (rendered_hand_pose_mt.py:99-147),
        # three gated 64x64 label-map views (real_animal_all_mt.py:275-283,306-311)
        multi_cache = {}

        def multi_bufs(r):
            if r not in multi_cache:
                multi_cache[r] = dict(big=[torch.empty(planes * 4096, dtype=torch.float32, device=dev) for _ in range(3)],
                                      small=[torch.empty(planes * 64, dtype=torch.float32, device=dev) for _ in range(2)],
                                      w=[torch.empty(planes, dtype=torch.float32, device=dev) for _ in range(5)],
                                      v=[torch.empty(planes, dtype=torch.int32, device=dev) for _ in range(3)])
            return multi_cache[r]
        gate = (vd > 0.5).to(torch.uint8).contiguous()

        def mk_gt_multi(r):
            d = multi_bufs(r)
            jobs = (_lib.TargetJob * 5)()
            outs = [d["big"][0], d["big"][1], d["small"][0], d["big"][2], d["small"][1]]
            for i, o in enumerate(outs):
                jobs[i].joints, jobs[i].vis, jobs[i].target, jobs[i].weight = jd.data_ptr(), vd.data_ptr(), o.data_ptr(), d["w"][i].data_ptr()
                jobs[i].hm_w = jobs[i].hm_h = 64 if o.numel() == planes * 4096 else 8
            return lambda: chk(lib.udape_gauss_target_multi(jobs, 5, planes, float(sigma), 256.0, 256.0, st()))

        def mk_lm_multi(r):
            d = multi_bufs(r)
            jobs = (_lib.LabelmapJob * 3)()
            for i in range(3):
                jobs[i].pts, jobs[i].gate, jobs[i].img, jobs[i].vis_out = pts.data_ptr(), gate.data_ptr(), d["big"][i].data_ptr(), d["v"][i].data_ptr()
            return lambda: chk(lib.udape_labelmap_multi(jobs, 3, planes, 64, 64, float(sigma), 0, st()))

        multi_bytes = 3 * hm32 + 2 * planes * 64 * 4 + 5 * 24 * planes
        bench("gauss_target x5 (1 launch)", shape, multi_bytes, mk_gt_multi, "target")
        bench("labelmap x3 (1 launch)", shape, 3 * hm32 + 3 * 13 * planes, mk_lm_multi, "target")
        # re-warp (three tF.affine stages composed): teacher fp32 forward, student fp16 forward + backward
        from uda_poseestimation_b200 import rewarp as RW
        t32 = RW.stage_table(RW.recon_stages(S.aug_params(b, seed=5), 4.0, b), 64, 64, torch.float32, None)[0].to(dev)
        t16 = RW.stage_table(RW.recon_stages(S.aug_params(b, seed=6), 4.0, b), 64, 64, torch.float16, torch.float16)[0].to(dev)
        in_arr = lambda t: (ctypes.c_void_p * 1)(t.data_ptr())  # noqa: E731

        def mk_rw32(r):
            d = B(r)
            ia, ta = in_arr(d["tea"]), in_arr(t32)
            return lambda: chk(lib.udape_rewarp_fwd(ia, ta, 1, 3, 0, _lib.F16, None, 0, None, b, k, 64, 64, _lib.F32,
                                                    d["rect"].data_ptr(), None, st()))

        def mk_rwdec(r):   # the step's teacher chain in one launch: re-warp + decode + conf_table + k-th value mask, no map written
            d = B(r)
            tk = _lib.ticket(dev)
            return lambda: chk(lib.udape_rewarp_decode_select(d["tea"].data_ptr(), t32.data_ptr(), 3, 0, _lib.F16, b, k, 64, 64, _lib.F32,
                                                              None, d["preds"].data_ptr(), d["maxv"].data_ptr(), d["pos"].data_ptr(), 0.9,
                                                              d["conf"].data_ptr(), planes // 2, None, d["thresh"].data_ptr(),
                                                              d["tm"].data_ptr(), tk, st()))

        def mk_rw16(r):
            d = B(r)
            ia, ta = in_arr(d["stu16"]), in_arr(t16)
            return lambda: chk(lib.udape_rewarp_fwd(ia, ta, 1, 3, 7, _lib.F16, None, 0, None, b, k, 64, 64, _lib.F16,
                                                    d["grad16"].data_ptr(), None, st()))

        plan16 = torch.empty(b, lib.udape_rewarp_plan_elems(64, 64, 2), dtype=torch.int16, device=dev)

        def mk_rw16p(r):   # forward that also writes the inverse plan (what autograd runs)
            d = B(r)
            ia, ta = in_arr(d["stu16"]), in_arr(t16)
            return lambda: chk(lib.udape_rewarp_fwd(ia, ta, 1, 3, 7, _lib.F16, None, 0, None, b, k, 64, 64, _lib.F16,
                                                    d["grad16"].data_ptr(), plan16.data_ptr(), st()))

        def mk_rwb(r):
            d = B(r)
            return lambda: chk(lib.udape_rewarp_bwd(d["stu16"].data_ptr(), t16.data_ptr(), 3, 7, _lib.F16, b, k, 64, 64, _lib.F16,
                                                    d["grad16"].data_ptr(), None, st()))

        def mk_rwbp(r):
            d = B(r)
            return lambda: chk(lib.udape_rewarp_bwd(d["stu16"].data_ptr(), t16.data_ptr(), 3, 7, _lib.F16, b, k, 64, 64, _lib.F16,
                                                    d["grad16"].data_ptr(), plan16.data_ptr(), st()))

        bench("rewarp_fwd f32 (teacher)", shape, 2 * hm32, mk_rw32, "rewarp")
        bench("rewarp f32 + decode + kth-select (1 launch)", shape, hm32 + 41 * planes, mk_rwdec, "rewarp")
        bench("rewarp_fwd f16 (student)", shape, 2 * hm16, mk_rw16, "rewarp")
        bench("rewarp_fwd f16 + inv. plan", shape, 2 * hm16 + plan16.numel() * 2, mk_rw16p, "rewarp")
        bench("rewarp_bwd f16 (no plan)", shape, 2 * hm16, mk_rwb, "rewarp")
        mk_rw16p(0)()   # a valid plan for the plan-based backward
        bench("rewarp_bwd f16 (plan)", shape, 2 * hm16 + plan16.numel() * 2, mk_rwbp, "rewarp")
        bundles.clear()

    # ---- decoder pre-training job: statistics of relu1_1..relu4_1 (N=8, 256x256 crops) + their backward ----
    for shp in ((8, 64, 256, 256), (8, 128, 128, 128), (8, 256, 64, 64), (8, 512, 32, 32)):
        n_, c_, h_, w_ = shp
        cache = {}

        def stat_bufs(r, shp=shp, cache=cache):
            if r not in cache:
                x = torch.relu(torch.randn(*shp, device=dev))
                pl = shp[0] * shp[1]
                cache[r] = (x, torch.empty(pl, device=dev), torch.empty(pl, device=dev), torch.randn(pl, device=dev),
                            torch.randn(pl, device=dev), torch.empty_like(x))
            return cache[r]

        def mk_stat(r, shp=shp, stat_bufs=stat_bufs):
            x, m, s_, _, _, _ = stat_bufs(r)
            return lambda: chk(lib.udape_mean_std(x.data_ptr(), _lib.F32, shp[0] * shp[1], shp[2] * shp[3], 1e-5,
                                                  m.data_ptr(), s_.data_ptr(), st()))

        def mk_stat_bwd(r, shp=shp, stat_bufs=stat_bufs):
            x, m, s_, dm, ds, dx = stat_bufs(r)
            chk(lib.udape_mean_std(x.data_ptr(), _lib.F32, shp[0] * shp[1], shp[2] * shp[3], 1e-5, m.data_ptr(), s_.data_ptr(), st()))
            return lambda: chk(lib.udape_mean_std_bwd(x.data_ptr(), m.data_ptr(), s_.data_ptr(), dm.data_ptr(), ds.data_ptr(),
                                                      _lib.F32, shp[0] * shp[1], shp[2] * shp[3], dx.data_ptr(), st()))

        nb = n_ * c_ * h_ * w_ * 4
        bench("mean_std (pre-training)", f"{n_}x{c_}x{h_}x{w_} f32", nb, mk_stat, "stats")
        bench("mean_std_bwd", f"{n_}x{c_}x{h_}x{w_} f32", 2 * nb, mk_stat_bwd, "stats")
        cache.clear()

    # ---- per-channel clamp of the stylised images (train_human.py:276) -----------------------------------
    for n in (32, 64):
        cache = {}

        def mk_clamp(r, n=n, cache=cache):
            if r not in cache:
                cache[r] = (torch.randn(n, 3, 256, 256, device=dev) * 2, torch.empty(n, 3, 256, 256, device=dev))
            x, o = cache[r]
            lo = torch.tensor([-2.1179, -2.0357, -1.8044], device=dev)
            hi = torch.tensor([2.2489, 2.4285, 2.64], device=dev)
            return lambda: chk(lib.udape_channel_clamp(x.data_ptr(), _lib.F32, n * 3, 3, 65536, lo.data_ptr(), hi.data_ptr(),
                                                       o.data_ptr(), st()))

        bench("channel_clamp", f"{n}x3x256x256 f32", 2 * n * 3 * 65536 * 4, mk_clamp, "clamp")
        cache.clear()

    # ---- EMA over the PoseResNet-101 parameter census -----------------------------------------------------
    shapes = S.pose_resnet_param_shapes(21)
    student = S.parameter_list(shapes, 1, device=dev)
    teacher = [t.clone() for t in student]
    plan = MultiTensorPlan(teacher, student)
    n_params = sum(t.numel() for t in student)
    bench("ema_multi f32", f"PoseResNet-101 {n_params}", 3 * n_params * 4, lambda r: (lambda: plan.run(0.999, 0.001, 0)), "ema")
    plan_c = MultiTensorPlan(teacher, student, as_bytes=True)
    bench("ema_multi copy", f"PoseResNet-101 {n_params}", 2 * n_params * 4, lambda r: (lambda: plan_c.run(0.0, 1.0, 1)), "ema")
    # what the reference runs for the same work on the same device (eager torch op sequences, written
    # out here from adain/function.py:3-22 + Style_net.py:167-168 and utils.py:21-25; wall time on the stream)
    if (not only or "ema" in only or "adain" in only) and not args.no_sustained:
        def eager_time(fn, reps=5):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record(); fn(); b_.record(); torch.cuda.synchronize()
                ts.append(a_.elapsed_time(b_))
            return float(np.median(ts)) * 1e3

        c_ = torch.relu(torch.randn(32, 512, 32, 32, device=dev))
        s_ = torch.relu(torch.randn(32, 512, 32, 32, device=dev) * 2)

        def eager_stats(feat, eps=1e-5):
            n_, ch = feat.size()[:2]
            var = feat.view(n_, ch, -1).var(dim=2) + eps
            return feat.view(n_, ch, -1).mean(dim=2).view(n_, ch, 1, 1), var.sqrt().view(n_, ch, 1, 1)

        def eager_adain_mix():
            size = c_.size()
            sm, ss = eager_stats(s_)
            cm, cs = eager_stats(c_)
            t = (c_ - cm.expand(size)) / cs.expand(size) * ss.expand(size) + sm.expand(size)
            return 0.37 * t + (1 - 0.37) * c_

        def eager_ema():
            with torch.no_grad():
                for tp, sp in zip(teacher, student):
                    tp.mul_(0.999)
                    tp.add_(sp * 0.001)

        for label, fn, nb in (("torch eager adain+mix (reference ops)", eager_adain_mix, 3 * c_.numel() * 4),
                              ("torch eager OldWeightEMA loop (969 launches)", eager_ema, 3 * n_params * 4)):
            us = eager_time(fn)
            rows.append(dict(kernel=label, shape="32x512x32x32 f32" if "adain" in label else f"PoseResNet-101 {n_params}",
                             mbytes=nb / 1e6, us=us, gbs=nb / us / 1e3, note="eager reference op sequence, not a udape kernel"))
            print(f"{label:<48}{nb / 1e6:9.1f} MB {us:8.1f} us {nb / us / 1e3:7.0f} GB/s (algorithmic bytes)", flush=True)
        del c_, s_

    # reference point: torch's own copy of the same bytes (what MEASURED_PEAKS measures)
    big_a = torch.empty(n_params, device=dev)
    big_b = torch.empty(n_params, device=dev)
    bench("torch copy_ (reference pt)", f"{n_params} f32", 2 * n_params * 4, lambda r: (lambda: big_b.copy_(big_a)), "ema")

    # ---- student step: unscale + Adam | SGD + teacher EMA in one launch (train_human.py:436-438) ---------
    if not only or "optim" in only:
        import uda_poseestimation_b200 as U

        class Bag(torch.nn.Module):
            def __init__(self, tensors):
                super().__init__()
                self.ps = torch.nn.ParameterList([torch.nn.Parameter(t) for t in tensors])

        stu, tea_m = Bag(student), Bag(teacher)
        for p_ in stu.parameters():
            p_.grad = torch.randn_like(p_) * 65.536
        scale_t = torch.full((), 65536.0, device=dev)
        for label, mk_opt, passes in (("adam+ema", lambda: U.Adam(stu.parameters(), lr=1e-3), 9),
                                      ("sgd-nesterov+ema", lambda: U.SGD(stu.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4,
                                                                         nesterov=True), 7)):
            opt = mk_opt()
            ema_o = U.OldWeightEMA(tea_m, stu, alpha=0.999)
            opt.attach_teacher(ema_o)
            opt.grad_scale, opt.found_inf = scale_t, opt.check_grads()
            opt.step()   # builds tables / state outside the timed (and captured) region

            def mk_step(r, opt=opt, ema_o=ema_o):
                def go():
                    opt.step()
                    ema_o.step()
                return go

            bench(f"student_step {label}", f"PoseResNet-101 {n_params}", passes * n_params * 4, mk_step, "optim")
            if label.startswith("adam"):
                bench("grad_check (found_inf)", f"PoseResNet-101 {n_params}", n_params * 4,
                      lambda r, opt=opt: (lambda: opt.check_grads()), "optim")
            del opt.grad_scale, opt.found_inf
            del opt, ema_o
        # what the reference runs on the same device: GradScaler.unscale_ + foreach Adam + per-tensor EMA
        ref_opt = torch.optim.Adam(stu.parameters(), lr=1e-3)
        ref_opt.step()
        inv = torch.full((), 1.0 / 65536.0, device=dev)
        finf = torch.zeros((), device=dev)
        grads_l = [p_.grad for p_ in stu.parameters()]
        tps, sps = list(tea_m.parameters()), list(stu.parameters())

        def ref_step():
            torch._amp_foreach_non_finite_check_and_unscale_(grads_l, finf, inv)
            ref_opt.step()
            with torch.no_grad():
                for tp, sp in zip(tps, sps):          # utils.py:21-25
                    tp.data.mul_(0.999)
                    tp.data.add_(sp.data * 0.001)

        if not args.no_sustained:
            for _ in range(2):
                ref_step()
            torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(5):
                a_.record(); ref_step(); b_.record(); torch.cuda.synchronize()
                ts.append(a_.elapsed_time(b_))
            t_ref = float(np.median(ts))
            rows.append(dict(kernel="torch eager: unscale_ + Adam(foreach) + OldWeightEMA loop", shape=f"PoseResNet-101 {n_params}",
                             mbytes=9 * n_params * 4 / 1e6, us=t_ref * 1e3, note="host-launched eager sequence, wall time on the stream"))
            print(f"{'torch eager unscale+Adam+EMA':<28}{'PoseResNet-101':<26}{9 * n_params * 4 / 1e6:9.1f} MB {t_ref * 1e3:8.1f} us", flush=True)

    out = Path(args.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(dict(peak_gbs=peak, device=torch.cuda.get_device_name(0), iters=args.iters,
                                   l2_flush=not args.no_flush, rows=rows), indent=1))
    print(f"wrote {out}")


if __name__ == "__main__":
    main()

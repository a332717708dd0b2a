"""Heatmap arg-max decoding and PCK accuracy — drop-in for the reference's
``lib/keypoint_detection.py`` (``get_max_preds`` :9-37, ``calc_dists`` :40-52, ``dist_acc``
:55-62, ``accuracy`` :65-94) and ``utils.py:54-75`` (``get_max_preds_torch``).

The reference decodes on the CPU in numpy after a device→host copy of the full heatmaps
(``train_human.py:443-444``) and counts hits in a B×K Python loop.  Here the planes are
decoded by one CUDA launch (``csrc/decode.cu``) and the PCK hit/valid counts are integer
atomics, so results are bit-identical to numpy/torch (first-index tie-break, NaN is max).

* numpy in → numpy out (the reference contract; the array is copied to the GPU and back);
* CUDA tensor in → the D2H copy of the heatmaps is avoided; ``accuracy`` still returns the
  reference's host tuple, ``pck_counts`` stays fully on the device (for the multi-GPU
  integer all-reduce).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ["get_max_preds", "get_max_preds_torch", "calc_dists", "dist_acc", "accuracy", "pck_counts",
           "accuracy_from_counts", "decode"]


def _as_cuda_heatmap(x, name: str):
    """numpy / torch → contiguous CUDA tensor; remembers whether the caller gave numpy."""
    if isinstance(x, np.ndarray):
        if x.ndim != 4:
            raise AssertionError("batch_images should be 4-ndim")  # keypoint_detection.py:16
        if x.dtype not in (np.float32, np.float16):
            raise TypeError(f"{name}: numpy dtype {x.dtype} is not supported (float32/float16 only)")
        if not torch.cuda.is_available():
            raise RuntimeError(f"{name}: no CUDA device; this package has no CPU fallback")
        return torch.from_numpy(np.ascontiguousarray(x)).cuda(non_blocking=False), True
    if not isinstance(x, torch.Tensor):
        raise AssertionError("batch_heatmaps should be numpy.ndarray")  # keypoint_detection.py:14
    if x.dim() != 4:
        raise AssertionError("batch_images should be 4-ndim")
    _lib.require_cuda(x)
    return x.detach().contiguous(), False


def decode(hm: torch.Tensor, *, want_idx=False, want_preds=False, want_maxvals=False,
           want_maxvals_f32=False, want_position=False, occlude_thresh: float | None = None,
           rectify_sigma: float | None = None, select_kth: int | None = None,
           select_tea_mask: torch.Tensor | None = None) -> dict:
    """One launch of ``udape_decode`` over ``hm[B,K,H,W]``; returns only the requested outputs.

    Keys: ``idx`` int32[B,K], ``preds`` float32[B,K,2], ``maxvals`` hm.dtype[B,K,1],
    ``maxvals_f32`` float32[B,K], ``position`` int64[B,K,2], ``conf_table`` bool[B,K],
    ``rectified`` hm.dtype[B,K,H,W].  With ``select_kth`` (1-based rank, train_human.py:429) the same
    launch also selects the k-th smallest activation (``udape_decode_select``): ``mask_thresh`` 0-dim
    float32 and ``tea_mask`` bool[B,K] = ``(select_tea_mask * maxvals) > mask_thresh``.
    """
    dev = _lib.require_cuda(hm)
    if hm.dim() != 4:
        raise AssertionError("batch_images should be 4-ndim")
    hm = hm.contiguous()
    b, k, h, w = hm.shape
    code = _lib.float_code(hm)
    out = {}
    planes = b * k
    idx = torch.empty((b, k), dtype=torch.int32, device=dev) if want_idx else None
    preds = torch.empty((b, k, 2), dtype=torch.float32, device=dev) if want_preds else None
    maxvals = torch.empty((b, k, 1), dtype=hm.dtype, device=dev) if want_maxvals else None
    mv32 = torch.empty((b, k), dtype=torch.float32, device=dev) if want_maxvals_f32 else None
    pos = torch.empty((b, k, 2), dtype=torch.int64, device=dev) if want_position else None
    conf = torch.empty((b, k), dtype=torch.bool, device=dev) if occlude_thresh is not None else None
    if occlude_thresh is not None and hm.dtype != torch.float32:
        # `conf >= args.occlude_thresh` on a half tensor compares in the tensor's dtype: torch rounds the Python
        # scalar to half first (a plane whose maximum is half(0.9) = 0.89990234 passes).  The kernel compares
        # the exactly-widened maximum in float32, so the threshold is rounded here
        occlude_thresh = float(torch.tensor(float(occlude_thresh), dtype=hm.dtype))
    rect = torch.empty_like(hm) if rectify_sigma is not None else None
    tea_mask = thresh = None
    if select_kth is not None:
        if not (1 <= select_kth <= planes):
            raise IndexError(f"kthvalue(): selected number k out of range for dimension 0 (k={select_kth}, n={planes})")
        if mv32 is None:
            mv32 = torch.empty((b, k), dtype=torch.float32, device=dev)
        tm_in = None
        if select_tea_mask is not None:
            _lib.require_cuda(hm, select_tea_mask)
            if select_tea_mask.numel() != planes:
                raise ValueError("decode: select_tea_mask must have one entry per (b, k)")
            tm_in = select_tea_mask.detach().to(torch.float32).contiguous()
        tea_mask = torch.empty((b, k), dtype=torch.bool, device=dev)
        thresh = torch.empty((), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            st = _lib.load().udape_decode_select(
                hm.data_ptr(), code, planes, h, w, _lib.ptr(idx), _lib.ptr(preds), _lib.ptr(maxvals),
                _lib.ptr(mv32), _lib.ptr(pos), float(occlude_thresh if occlude_thresh is not None else 0.0),
                _lib.ptr(conf), float(rectify_sigma if rectify_sigma is not None else 1.0), _lib.ptr(rect),
                int(select_kth), _lib.ptr(tm_in), thresh.data_ptr(), tea_mask.data_ptr(), _lib.ticket(dev),
                _lib.stream_ptr(dev))
        _lib.check(st, "decode")
        out["tea_mask"], out["mask_thresh"] = tea_mask, thresh
        if want_maxvals_f32:
            out["maxvals_f32"] = mv32
    elif planes > 0 and h * w > 0:
        with _lib.on_device(dev):
            st = _lib.load().udape_decode(
                hm.data_ptr(), code, planes, h, w, _lib.ptr(idx), _lib.ptr(preds), _lib.ptr(maxvals),
                _lib.ptr(mv32), _lib.ptr(pos), float(occlude_thresh if occlude_thresh is not None else 0.0),
                _lib.ptr(conf), float(rectify_sigma if rectify_sigma is not None else 1.0), _lib.ptr(rect),
                _lib.stream_ptr(dev))
        _lib.check(st, "decode")
    for key, val in (("idx", idx), ("preds", preds), ("maxvals", maxvals), ("maxvals_f32", mv32 if want_maxvals_f32 else None),
                     ("position", pos), ("conf_table", conf), ("rectified", rect)):
        if val is not None:
            out[key] = val
    return out


def get_max_preds(batch_heatmaps):
    """``(preds[B,K,2] float32, maxvals[B,K,1])`` from score maps (keypoint_detection.py:9-37).

    numpy input returns numpy (the reference contract); a CUDA tensor returns CUDA tensors.
    """
    hm, was_numpy = _as_cuda_heatmap(batch_heatmaps, "get_max_preds")
    r = decode(hm, want_preds=True, want_maxvals=True)
    if was_numpy:
        return r["preds"].cpu().numpy(), r["maxvals"].cpu().numpy()
    return r["preds"], r["maxvals"]


def get_max_preds_torch(batch_heatmaps: torch.Tensor):
    """Tensor version used by ``rectify`` (utils.py:54-75): same values, torch in / torch out."""
    if not isinstance(batch_heatmaps, torch.Tensor):
        raise TypeError("get_max_preds_torch expects a torch.Tensor")
    r = decode(batch_heatmaps.detach(), want_preds=True, want_maxvals=True)
    return r["preds"], r["maxvals"]


# ---- host-side helpers on decoded coordinates (tiny [B,K,2] arrays) -----------------------------
def calc_dists(preds, target, normalize):
    """Normalised float64 distances ``[K,B]`` between decoded coordinates, ``-1`` where the
    target is not strictly inside (x>1 and y>1) — keypoint_detection.py:40-52, vectorised.
    Host utility on [B,K,2] coordinate arrays; the device path is :func:`pck_counts`."""
    preds = np.asarray(preds).astype(np.float32)
    target = np.asarray(target).astype(np.float32)
    normalize = np.asarray(normalize, dtype=np.float64)
    d = preds / normalize[:, None, :] - target / normalize[:, None, :]
    dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
    ok = (target[..., 0] > 1) & (target[..., 1] > 1)
    return np.where(ok, dist, -1.0).T.copy()


def dist_acc(dists, thr=0.5):
    """Fraction of valid (``!= -1``) distances below ``thr``; ``-1`` if none (:55-62)."""
    dists = np.asarray(dists)
    valid = dists != -1
    n = int(valid.sum())
    if n == 0:
        return -1
    return int((dists[valid] < thr).sum()) * 1.0 / n


def _pck(output: torch.Tensor, target: torch.Tensor, thr: float):
    """``(counts int32[2,K] = hits ‖ valid, pred float32[B,K,2])`` on the device."""
    dev = _lib.require_cuda(output, target)
    if output.dim() != 4 or output.shape != target.shape:
        raise AssertionError(f"accuracy: output {tuple(output.shape)} and target {tuple(target.shape)} must be equal 4-D shapes")
    output = output.detach().contiguous()
    target = target.detach().contiguous()
    b, k, h, w = output.shape
    counts = torch.empty((2, k), dtype=torch.int32, device=dev)
    pred = torch.empty((b, k, 2), dtype=torch.float32, device=dev)
    if b * k == 0:
        counts.zero_()
        return counts, pred
    flags = torch.empty(b * k, dtype=torch.uint8, device=dev)  # per-(b,k) valid/hit bits (scratch)
    with _lib.on_device(dev):
        st = _lib.load().udape_pck_counts(output.data_ptr(), _lib.float_code(output), target.data_ptr(),
                                          _lib.float_code(target), b, k, h, w, float(thr), pred.data_ptr(),
                                          None, counts[0].data_ptr(), counts[1].data_ptr(), flags.data_ptr(),
                                          _lib.ticket(dev), _lib.stream_ptr(dev))
    _lib.check(st, "accuracy")
    return counts, pred


def pck_counts(output: torch.Tensor, target: torch.Tensor, thr: float = 0.5):
    """Device-side PCK: ``(hits int32[K], valid int32[K], pred float32[B,K,2])``, no host sync.

    ``hits`` and ``valid`` are the two rows of one contiguous int32[2,K] tensor
    (``hits._base``), which is what the multi-GPU path all-reduces before
    the per-joint ratios of ``accuracy`` are formed.
    """
    counts, pred = _pck(output, target, thr)
    return counts[0], counts[1], pred


def accuracy_from_counts(hits, valid):
    """``(acc float64[K], avg_acc, cnt)`` from integer counts, as keypoint_detection.py:82-92."""
    hits = np.asarray(hits.cpu() if isinstance(hits, torch.Tensor) else hits, dtype=np.int64)
    valid = np.asarray(valid.cpu() if isinstance(valid, torch.Tensor) else valid, dtype=np.int64)
    acc = np.zeros(len(hits))
    avg_acc = 0
    cnt = 0
    for i in range(len(hits)):
        # dist_acc: hits * 1.0 / valid, or -1 when the joint has no valid sample
        acc[i] = hits[i] * 1.0 / valid[i] if valid[i] > 0 else -1
        if acc[i] >= 0:
            avg_acc = avg_acc + acc[i]
            cnt += 1
    avg_acc = avg_acc / cnt if cnt != 0 else 0
    return acc, avg_acc, cnt


def accuracy(output, target, hm_type="gaussian", thr=0.5):
    """PCK from predicted and ground-truth heatmaps — ``(acc[K], avg_acc, cnt, pred[B,K,2])``
    exactly as keypoint_detection.py:65-94 (numpy outputs; one small D2H of 2K ints + pred)."""
    if hm_type != "gaussian":
        raise ValueError("accuracy: only hm_type='gaussian' is defined by the reference (:73-79)")
    out_t, _ = _as_cuda_heatmap(output, "accuracy")
    tgt_t, _ = _as_cuda_heatmap(target, "accuracy")
    counts, pred = _pck(out_t, tgt_t, thr)
    counts = counts.cpu().numpy()
    acc, avg_acc, cnt = accuracy_from_counts(counts[0], counts[1])
    return acc, avg_acc, cnt, pred.cpu().numpy()

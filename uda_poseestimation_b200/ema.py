"""EMA teacher update — drop-in for ``utils.py:9-25`` (``OldWeightEMA``, the optimizer the
trainers actually use, ``train_human.py:141,438``) and ``lib/models/ema.py:6-44``
(``ModelEMA``, API mirrored for completeness).

The reference updates one parameter tensor at a time (``mul_``, a temporary, ``add_``:
969 launches and 7 passes over 212 MB for PoseResNet-101).  Here all tensors are described
once by a device-resident chunk table and updated by ONE launch of ``udape_ema_multi``
(2 reads + 1 write per element).  The live ``Parameter`` storages are updated in place, so
``state_dict()`` / checkpoints / DataParallel wrappers keep seeing the same tensors.  The
fp32 arithmetic ``fl(fl(p*a) + fl(s*(1-a)))`` is bit-identical to the eager reference.
"""
from __future__ import annotations

import ctypes
import os
from copy import deepcopy

import torch

from . import _lib

__all__ = ["OldWeightEMA", "ModelEMA", "MultiTensorPlan"]

# elements per CTA: 256 threads x 4 x 128-bit vectors (fp32).  UDAPE_EMA_CHUNK overrides it (tuning).
CHUNK_ELEMS = int(os.environ.get("UDAPE_EMA_CHUNK", "4096"))


def _dense_like(p: torch.Tensor, s: torch.Tensor) -> bool:
    if p.shape != s.shape:
        return False
    if p.is_contiguous() and s.is_contiguous():
        return True
    # same dense permuted layout on both sides (e.g. channels_last conv weights): the update
    # is elementwise, so walking the two storages linearly is still element-aligned
    return (p.stride() == s.stride() and p.dim() == 4
            and p.is_contiguous(memory_format=torch.channels_last))


class MultiTensorPlan:
    """Device chunk tables for a list of (dst, src) tensor pairs, grouped by device/dtype."""

    def __init__(self, dst, src, *, as_bytes: bool = False, chunk_elems: int = CHUNK_ELEMS):
        dst, src = list(dst), list(src)
        if len(dst) != len(src):
            raise ValueError("MultiTensorPlan: dst/src length mismatch")
        self.dst, self.src = dst, src
        self.as_bytes = as_bytes
        self.chunk_elems = chunk_elems * (4 if as_bytes else 1)
        self.groups = []  # (device, dtype_code, n_chunks, table_tensor)
        self.numel = 0
        buckets = {}
        for d, s in zip(dst, src):
            dev = _lib.require_cuda(d, s)
            if d.dtype != s.dtype:
                raise TypeError(f"EMA: dtype mismatch {d.dtype} vs {s.dtype}")
            if not _dense_like(d, s):
                raise ValueError(f"EMA: tensors must be dense with identical layout, got shapes/strides "
                                 f"{tuple(d.shape)}/{d.stride()} vs {tuple(s.shape)}/{s.stride()}")
            if d.numel() == 0:
                continue
            if as_bytes:
                key = (dev, _lib.U8, 1)
                n = d.numel() * d.element_size()
            else:
                key = (dev, _lib.float_code(d), d.element_size())
                n = d.numel()
            buckets.setdefault(key, []).append((d.data_ptr(), s.data_ptr(), n))
            self.numel += d.numel()
        lib = _lib.load()
        for (dev, code, esize), items in buckets.items():
            n_t = len(items)
            dptr = (ctypes.c_void_p * n_t)(*[it[0] for it in items])
            sptr = (ctypes.c_void_p * n_t)(*[it[1] for it in items])
            numel = (ctypes.c_int64 * n_t)(*[it[2] for it in items])
            need = lib.udape_ema_plan(dptr, sptr, numel, n_t, esize, self.chunk_elems, None, 0)
            if need < 0:
                _lib.check(int(need), "udape_ema_plan")
            table = (_lib.EmaChunk * need)()
            got = lib.udape_ema_plan(dptr, sptr, numel, n_t, esize, self.chunk_elems, table, need)
            if got != need:
                _lib.check(int(got) if got < 0 else -3, "udape_ema_plan")
            host = torch.frombuffer(table, dtype=torch.uint8).clone()
            self.groups.append((dev, code, int(need), host.to(dev)))
        self._ptrs = self.pointer_signature()

    def pointer_signature(self):
        return tuple(t.data_ptr() for t in self.dst) + tuple(t.data_ptr() for t in self.src)

    def stale(self) -> bool:
        """True if any tensor was re-allocated since the plan was built."""
        return self.pointer_signature() != self._ptrs

    def run(self, a: float, b: float, mode: int) -> None:
        lib = _lib.load()
        for dev, code, n_chunks, table in self.groups:
            with _lib.on_device(dev):
                st = lib.udape_ema_multi(table.data_ptr(), n_chunks, self.chunk_elems, a, b, code, mode,
                                         _lib.stream_ptr(dev))
            _lib.check(st, "udape_ema_multi")


class OldWeightEMA(object):
    """Exponential-moving-average weight "optimizer" for the mean-teacher model
    (utils.py:9-25): ``__init__`` copies source → target, ``step()`` does
    ``p = p * alpha + src * (1 - alpha)`` over ``parameters()`` (buffers untouched)."""

    def __init__(self, target_net, source_net, alpha=0.999):
        self.target_params = list(target_net.parameters())
        self.source_params = list(source_net.parameters())
        self.alpha = alpha
        self._plan = None
        self._fused_pending = False  # set by optim.Adam/SGD.step() when it already folded this EMA in
        n = min(len(self.target_params), len(self.source_params))  # zip() semantics
        dst = [p.data for p in self.target_params[:n]]
        src = [p.data for p in self.source_params[:n]]
        if n and all(t.is_cuda for t in dst + src):
            MultiTensorPlan(dst, src, as_bytes=True).run(0.0, 1.0, 1)  # p.data[:] = src_p.data[:]
        else:
            # the trainers build this object BEFORE the models move to the GPU (train_human.py:141 vs
            # :145-146): the one-off initial copy of utils.py:18-19 is then a host copy; step() is the
            # CUDA operator and plans lazily, on the storages the parameters have by then
            for d, s in zip(dst, src):
                d.copy_(s)

    def _get_plan(self) -> MultiTensorPlan:
        plan = self._plan
        if plan is None or plan.stale():
            n = min(len(self.target_params), len(self.source_params))
            plan = MultiTensorPlan([p.data for p in self.target_params[:n]],
                                   [p.data for p in self.source_params[:n]])
            # signature is taken on the Parameter objects so re-pointed `.data` is noticed
            plan.dst = self.target_params[:n]
            plan.src = self.source_params[:n]
            plan._ptrs = plan.pointer_signature()
            self._plan = plan
        return plan

    def step(self):
        if self._fused_pending:
            # the student optimizer this EMA is attached to (optim.attach_teacher) applied it inside its
            # own launch, with the updated student, exactly once for this step
            self._fused_pending = False
            return
        one_minus_alpha = 1.0 - self.alpha
        self._get_plan().run(float(self.alpha), float(one_minus_alpha), 0)


class ModelEMA(object):
    """``lib/models/ema.py:6-44``: keeps a deep copy of ``model`` on the GPU; ``update`` does
    the EMA over parameters (handling a ``module.`` prefix) and copies buffers;
    ``momentum_update`` is the same EMA over ``parameters()`` with momentum ``m``."""

    def __init__(self, model, decay):
        self.ema = deepcopy(model)
        self.ema.cuda()
        self.decay = decay
        self.ema_has_module = hasattr(self.ema, "module")
        self.param_keys = [k for k, _ in self.ema.named_parameters()]
        self.buffer_keys = [k for k, _ in self.ema.named_buffers()]
        for p in self.ema.parameters():
            p.requires_grad_(False)
        self._plans = {}

    def _plans_for(self, model):
        key = id(model)
        entry = self._plans.get(key)
        if entry is not None and not entry[0].stale() and not (entry[1] is not None and entry[1].stale()):
            return entry
        needs_module = hasattr(model, "module") and not self.ema_has_module
        prefix = "module." if needs_module else ""
        m_params = dict(model.named_parameters())
        m_bufs = dict(model.named_buffers())
        e_params = dict(self.ema.named_parameters())
        e_bufs = dict(self.ema.named_buffers())
        pd = [e_params[k] for k in self.param_keys]
        ps = [m_params[prefix + k] for k in self.param_keys]
        pplan = MultiTensorPlan([t.data for t in pd], [t.data for t in ps])
        pplan.dst, pplan.src = pd, ps
        pplan._ptrs = pplan.pointer_signature()
        bplan = None
        if self.buffer_keys:
            bd = [e_bufs[k] for k in self.buffer_keys]
            bs = [m_bufs[prefix + k] for k in self.buffer_keys]
            bplan = MultiTensorPlan(bd, bs, as_bytes=True)
        entry = (pplan, bplan)
        self._plans[key] = entry
        return entry

    def update(self, model):
        pplan, bplan = self._plans_for(model)
        pplan.run(float(self.decay), float(1.0 - self.decay), 0)  # ema.py:30
        if bplan is not None:
            bplan.run(0.0, 1.0, 1)  # ema.py:31-36

    def momentum_update(self, model, m):
        dst = [p.data for p in self.ema.parameters()]
        src = [p.data for p in model.parameters()]
        n = min(len(dst), len(src))
        MultiTensorPlan(dst[:n], src[:n]).run(float(m), float(1 - m), 0)  # ema.py:43-44

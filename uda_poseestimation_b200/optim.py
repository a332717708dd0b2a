"""Student optimizer step fused with the teacher EMA — drop-ins for the three objects the reference's
train step drives at ``train_human.py:436-440``::

    scaler.scale(loss_all).backward()
    scaler.step(stu_optimizer)        # GradScaler.unscale_ + Adam | SGD(momentum, nesterov)   (:136-139)
    tea_optimizer.step()              # OldWeightEMA                                            (:141)
    scaler.update()

``Adam`` / ``SGD`` subclass ``torch.optim.Optimizer`` (same constructor arguments, ``param_groups`` for
``MultiStepLR``, ``state_dict`` with torch's keys), ``GradScaler`` subclasses ``torch.amp.GradScaler``.
Swapping the three constructors keeps the trainer's call sequence unchanged; what runs is

* ``udape_grad_check``   — one read-only pass over the gradients -> ``found_inf`` on the device (torch
  reads AND rewrites every gradient to unscale it and then syncs the host with ``.item()``);
* ``udape_student_step`` — ONE multi-tensor launch that unscales in registers, applies the update and,
  once ``attach_teacher(tea_optimizer)`` was called, folds the new student into the teacher EMA
  (Adam: 9 tensor passes instead of ~20; the following ``tea_optimizer.step()`` is then a no-op).

A step with a non-finite gradient leaves student, optimizer state and step count untouched and still runs
the EMA — what ``scaler.step`` + ``tea_optimizer.step`` do.  float32 parameters only (the trainers keep
fp32 master weights under autocast).  There is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .ema import CHUNK_ELEMS, OldWeightEMA

__all__ = ["Adam", "SGD", "GradScaler"]


class _FusedStudentOptimizer(torch.optim.Optimizer):
    """Shared machinery: chunk tables per param group, device step counter, GradScaler protocol."""

    # torch.amp.GradScaler.step(): hand grad_scale / found_inf to step() instead of unscaling itself
    _step_supports_amp_scaling = True
    _algo = _lib.OPT_ADAM

    def __init__(self, params, defaults, capturable: bool = False):
        super().__init__(params, defaults)
        self.capturable = capturable
        self._teacher: OldWeightEMA | None = None
        self._plans = None          # [(group, table tensor, n_chunks)], one entry per param group
        self._sig = None
        self._step_dev = None       # int32 device counter of applied updates
        self._lr_dev = {}
        self._ws = None             # grad_check workspace (ticket + OR word)
        self._found_inf = None
        # SGD: one int32 word per parameter (all groups, in order), non-zero while its momentum buffer has
        # never been written: torch clones the gradient into the buffer the first time THAT parameter is
        # updated (torch/optim/sgd.py `if buf is None`), and the buffers of a loaded checkpoint are never fresh
        self._fresh = None
        self._fresh_index = {}

    # -- teacher -----------------------------------------------------------------------------------
    def attach_teacher(self, tea_optimizer: OldWeightEMA):
        """Fold ``tea_optimizer``'s EMA into this optimizer's step; the next ``tea_optimizer.step()``
        after every ``step()`` becomes a no-op (the reference's call order is kept)."""
        if not isinstance(tea_optimizer, OldWeightEMA):
            raise TypeError("attach_teacher expects the package's OldWeightEMA")
        self._teacher = tea_optimizer
        self._plans = None
        return self

    def detach_teacher(self):
        """Stop folding the EMA into this optimizer's step (``tea_optimizer.step()`` is the plain EMA launch
        again).  The reference's ``pretrain()`` phase (train_human.py:243-300) steps the student optimizer
        WITHOUT ``tea_optimizer.step()``: attach for ``train()`` only, or detach around ``pretrain()``."""
        if self._teacher is not None:
            self._teacher._fused_pending = False
        self._teacher = None
        self._plans = None
        return self

    # -- tables ------------------------------------------------------------------------------------
    def _device(self) -> torch.device:
        for g in self.param_groups:
            for p in g["params"]:
                return _lib.require_cuda(p)
        raise ValueError("optimizer has no parameters")

    def _init_state(self, p, group):
        raise NotImplementedError

    def _signature(self):
        sig = []
        for g in self.param_groups:
            for p in g["params"]:
                sig.append(p.data_ptr())
                sig.append(p.grad.data_ptr() if p.grad is not None else 0)
        if self._teacher is not None:
            sig.extend(t.data_ptr() for t in self._teacher.target_params)
            sig.extend(s.data_ptr() for s in self._teacher.source_params)
        return tuple(sig)

    def _build(self):
        dev = self._device()
        lib = _lib.load()
        pair = {}
        if self._teacher is not None:
            for t, s in zip(self._teacher.target_params, self._teacher.source_params):
                pair[id(s)] = t
        claimed = set()
        plans = []
        n_all = sum(len(g["params"]) for g in self.param_groups)
        if self._fresh is None or self._fresh.numel() != n_all or self._fresh.device != dev:
            old, old_index = self._fresh, self._fresh_index
            self._fresh = torch.zeros(max(n_all, 1), dtype=torch.int32, device=dev)
            self._fresh_index, k = {}, 0
            for g_ in self.param_groups:
                for p in g_["params"]:
                    self._fresh_index[id(p)] = k
                    if old is not None and id(p) in old_index:
                        self._fresh[k] = old[old_index[id(p)]]     # add_param_group: keep what is known
                    k += 1
        fresh_base = 0
        for gi, group in enumerate(self.param_groups):
            rows = []
            group_fresh = (fresh_base, len(group["params"]))
            fresh_base += len(group["params"])
            for p in group["params"]:
                if _lib.require_cuda(p) != dev:
                    raise RuntimeError("all parameters of a fused optimizer must live on one device")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("fused student step: parameters must be contiguous float32 "
                                    f"(got {p.dtype}, contiguous={p.is_contiguous()})")
                g = p.grad
                if g is not None:
                    if g.is_sparse or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device:
                        raise TypeError("fused student step: gradients must be dense contiguous float32 on the parameter's device")
                    s1, s2 = self._init_state(p, group)
                else:
                    s1 = s2 = None
                t = pair.get(id(p))
                if t is not None:
                    if t.shape != p.shape or t.dtype != p.dtype or not t.is_contiguous() or t.device != p.device:
                        raise ValueError("teacher / student parameter mismatch")
                    claimed.add(id(p))
                if g is None and t is None:
                    continue
                fresh = None
                if self._algo == _lib.OPT_SGD and s1 is not None:
                    fresh = self._fresh.data_ptr() + 4 * self._fresh_index[id(p)]
                rows.append((p.data_ptr(), g.data_ptr() if g is not None else None,
                             s1.data_ptr() if s1 is not None else None, s2.data_ptr() if s2 is not None else None,
                             t.data_ptr() if t is not None else None, p.numel(), fresh))
            if gi == 0 and self._teacher is not None:
                # EMA pairs whose student parameter this optimizer does not own (frozen layers): EMA only
                for t, s in zip(self._teacher.target_params, self._teacher.source_params):
                    if id(s) not in claimed and not any(s is q for gr in self.param_groups for q in gr["params"]):
                        _lib.require_cuda(t, s)
                        if s.dtype != torch.float32 or not (s.is_contiguous() and t.is_contiguous()):
                            raise TypeError("fused student step: EMA-only pairs must be contiguous float32")
                        rows.append((s.data_ptr(), None, None, None, t.data_ptr(), s.numel(), None))
            rows = [r for r in rows if r[5] > 0]
            n_t = len(rows)
            if n_t == 0:
                plans.append((group, None, 0, group_fresh))
                continue
            cols = [(ctypes.c_void_p * n_t)(*[r[c] for r in rows]) for c in range(5)]
            cols.append((ctypes.c_void_p * n_t)(*[r[6] for r in rows]))
            numel = (ctypes.c_int64 * n_t)(*[r[5] for r in rows])
            need = lib.udape_opt_plan(*cols, numel, n_t, CHUNK_ELEMS, None, 0)
            if need < 0:
                _lib.check(int(need), "udape_opt_plan")
            table = (_lib.OptChunk * need)()
            got = lib.udape_opt_plan(*cols, numel, n_t, CHUNK_ELEMS, table, need)
            if got != need:
                _lib.check(int(got) if got < 0 else -3, "udape_opt_plan")
            host = torch.frombuffer(table, dtype=torch.uint8).clone()
            plans.append((group, host.to(dev), int(need), group_fresh))
        if self._step_dev is None:
            self._step_dev = torch.zeros((), dtype=torch.int32, device=dev)
            self._ws = torch.zeros(2, dtype=torch.int32, device=dev)
            self._found_inf = torch.zeros((), dtype=torch.float32, device=dev)
        self._plans = plans
        self._sig = self._signature()

    def _tables(self):
        self._device()  # CPU parameters fail here with the package's "CUDA-only" error
        capturing = torch.cuda.is_current_stream_capturing()
        if self._plans is None or (not capturing and self._signature() != self._sig):
            if capturing:
                raise RuntimeError("fused student step: run one eager step before CUDA-graph capture (it builds the chunk tables)")
            self._build()
        return self._plans

    # -- GradScaler support ----------------------------------------------------------------------------
    def check_grads(self) -> torch.Tensor:
        """``found_inf`` (device float32 scalar, 1.0 if any gradient element is non-finite) from one
        read-only pass — no unscaled copy is written and the host is not synchronised."""
        plans = self._tables()
        dev = self._step_dev.device
        lib = _lib.load()
        acc = None
        for i, (_, table, n, _f) in enumerate(plans):
            out = self._found_inf if i == 0 else torch.zeros((), dtype=torch.float32, device=dev)
            with _lib.on_device(dev):
                st = lib.udape_grad_check(_lib.ptr(table), n, out.data_ptr(), self._ws.data_ptr(), _lib.stream_ptr(dev))
            _lib.check(st, "udape_grad_check")
            acc = out if acc is None else torch.maximum(acc, out)
        return acc

    def sync_lr(self):
        """capturable=True: push the groups' current ``lr`` to device memory (call after a scheduler
        step, outside the captured graph)."""
        for gi, group in enumerate(self.param_groups):
            t = self._lr_dev.get(gi)
            if t is None:
                t = self._lr_dev[gi] = torch.empty((), dtype=torch.float32, device=self._device())
            t.fill_(float(group["lr"]))

    def _hyper(self, group) -> _lib.OptHyper:
        raise NotImplementedError

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        plans = self._tables()
        dev = self._step_dev.device
        lib = _lib.load()
        # set by torch.amp.GradScaler.step() for optimizers with _step_supports_amp_scaling
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        for t in (grad_scale, found_inf):
            if t is not None and (t.device != dev or t.dtype != torch.float32):
                raise RuntimeError("grad_scale / found_inf must be float32 tensors on the parameters' device")
        ema_a, ema_b = 1.0, 0.0
        if self._teacher is not None:
            ema_a = float(self._teacher.alpha)
            ema_b = float(1.0 - self._teacher.alpha)      # utils.py:22
        if self.capturable and len(self._lr_dev) != len(self.param_groups):
            self.sync_lr()
        last = max((i for i, p_ in enumerate(plans) if p_[2]), default=-1)
        for gi, (group, table, n, (f0, nf)) in enumerate(plans):
            if n == 0:
                continue
            h = self._hyper(group)
            h.ema_a, h.ema_b, h.step = ema_a, ema_b, 1
            lr_dev = self._lr_dev[gi].data_ptr() if self.capturable else None
            # every group reads the device step counter; only the last launch of the step advances it.  SGD:
            # each launch clears the "momentum buffer not written yet" words of its own group
            sgd_flags = self._algo == _lib.OPT_SGD and nf > 0
            with _lib.on_device(dev):
                st = lib.udape_student_step(table.data_ptr(), n, self._algo, ctypes.byref(h), lr_dev,
                                            _lib.ptr(grad_scale), _lib.ptr(found_inf), self._step_dev.data_ptr(),
                                            1 if gi == last else 0,
                                            self._fresh.data_ptr() + 4 * f0 if sgd_flags else None,
                                            nf if sgd_flags else 0,
                                            _lib.ticket(dev) if (gi == last or sgd_flags) else None,
                                            _lib.stream_ptr(dev))
            _lib.check(st, "udape_student_step")
        if self._teacher is not None:
            self._teacher._fused_pending = True
        return loss

    def zero_grad(self, set_to_none: bool = False):
        """Zeroes the gradients IN PLACE by default (torch's default frees them): the chunk tables
        hold the gradient addresses, so stable storage avoids re-planning every step and is what
        CUDA-graph capture needs.  ``set_to_none=True`` restores torch's behaviour."""
        return super().zero_grad(set_to_none=set_to_none)

    # -- checkpoints -------------------------------------------------------------------------------
    def applied_steps(self) -> int:
        """Number of updates applied so far (host sync)."""
        return int(self._step_dev.item()) if self._step_dev is not None else 0

    def state_dict(self):
        step = float(self.applied_steps())
        for st in self.state.values():
            if st:
                st["step"] = torch.tensor(step)
        sd = super().state_dict()
        if self._algo == _lib.OPT_SGD and self._fresh is not None:
            # a momentum buffer that was never written does not exist for torch.optim.SGD: drop it, so that the
            # first update after a reload clones the gradient (and the file reads like torch's own)
            fresh = self._fresh.cpu().tolist()
            k = 0
            for g in self.param_groups:
                for p in g["params"]:
                    if fresh[self._fresh_index[id(p)]] and k in sd["state"]:
                        sd["state"][k] = {n: v for n, v in sd["state"][k].items() if n != "momentum_buffer"}
                    k += 1
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [float(st["step"]) for st in self.state.values() if "step" in st]
        self._plans = None
        # every momentum buffer that came with the checkpoint has been written (torch.optim.SGD files carry
        # no 'step'); parameters without one get their flag when _init_state creates the buffer
        self._fresh = None
        self._fresh_index = {}
        if steps:
            dev = self._device()
            if self._step_dev is None:
                self._step_dev = torch.zeros((), dtype=torch.int32, device=dev)
                self._ws = torch.zeros(2, dtype=torch.int32, device=dev)
                self._found_inf = torch.zeros((), dtype=torch.float32, device=dev)
            self._step_dev.fill_(int(max(steps)))


class Adam(_FusedStudentOptimizer):
    """``torch.optim.Adam(params, lr, betas, eps, weight_decay)`` (train_human.py:139) as one launch."""

    _algo = _lib.OPT_ADAM

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, *, capturable=False):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay), capturable)

    def _init_state(self, p, group):
        st = self.state[p]
        if "exp_avg" not in st:
            st["step"] = torch.tensor(0.0)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return st["exp_avg"], st["exp_avg_sq"]

    def _hyper(self, group):
        h = _lib.OptHyper()
        h.lr, (h.beta1, h.beta2) = float(group["lr"]), group["betas"]
        h.eps, h.weight_decay, h.nesterov = float(group["eps"]), float(group["weight_decay"]), 0
        return h


class SGD(_FusedStudentOptimizer):
    """``torch.optim.SGD(params, lr, momentum, dampening, weight_decay, nesterov)`` (train_human.py:137)."""

    _algo = _lib.OPT_SGD

    def __init__(self, params, lr=1e-3, momentum=0, dampening=0, weight_decay=0, nesterov=False, *, capturable=False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if momentum < 0.0:
            raise ValueError(f"Invalid momentum value: {momentum}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov), capturable)

    def _init_state(self, p, group):
        if group["momentum"] == 0:
            return None, None
        st = self.state[p]
        if st.get("momentum_buffer") is None:
            st.setdefault("step", torch.tensor(0.0))
            st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            self._fresh[self._fresh_index[id(p)]] = 1   # torch: `buf = torch.clone(grad)` on its first update
        return st["momentum_buffer"], None

    def _hyper(self, group):
        h = _lib.OptHyper()
        h.lr, h.beta1, h.beta2 = float(group["lr"]), float(group["momentum"]), float(group["dampening"])
        h.eps, h.weight_decay, h.nesterov = 0.0, float(group["weight_decay"]), int(bool(group["nesterov"]))
        return h


class GradScaler(torch.amp.GradScaler):
    """``torch.cuda.amp.GradScaler()`` (train_human.py:260,324) whose non-finite check of a fused student
    optimizer is the read-only ``udape_grad_check`` pass (no gradient rewrite, no ``.item()`` sync);
    every other optimizer takes torch's path unchanged."""

    def __init__(self, device: str = "cuda", **kwargs):
        super().__init__(device, **kwargs)

    def _check_inf_per_device(self, optimizer):
        from .dp import ShardedStudentStep
        if isinstance(optimizer, ShardedStudentStep):
            # the data-parallel step checks the REDUCED gradient inside step() (between its reduce-scatter and
            # the update); update() reads this same device scalar afterwards
            state = self._per_optimizer_states[id(optimizer)]
            state["found_inf_per_device"] = {optimizer.found_inf.device: optimizer.found_inf}
            return state["found_inf_per_device"]
        if isinstance(optimizer, _FusedStudentOptimizer):
            found = optimizer.check_grads()
            state = self._per_optimizer_states[id(optimizer)]
            state["found_inf_per_device"] = {found.device: found}
            return state["found_inf_per_device"]
        return super()._check_inf_per_device(optimizer)

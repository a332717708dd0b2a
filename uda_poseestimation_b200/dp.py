"""Data-parallel tail of the train step over peer memory — one process per GPU instead of the reference's
single-process ``nn.DataParallel`` (``train_human.py:145-148``) followed by
``scaler.step(stu_optimizer); tea_optimizer.step()`` (``:436-438``).

``PeerGroup`` gives every rank one zeroed, ``cudaMalloc``'ed arena that all other ranks of the node have
mapped (CUDA IPC, exchanged once through ``torch.distributed``); ``ShardedStudentStep`` lays the student's
parameters and gradients out flat inside it and runs the step as the three peer-memory kernels of
``csrc/dp.cu`` (``include/udape.h`` section e):

    reduce-scatter of the gradient buckets fused with GradScaler's non-finite check and unscale + Adam | SGD on
       this rank's 1/W slice (optimizer state is sharded; the update is speculative, into a shadow slice)
    -> all-gather of the shadow slices (the commit, unless any rank saw a non-finite gradient) fused with the
       teacher EMA (``OldWeightEMA``, utils.py:21-25)

with NVLink loads between the GPUs and no NCCL call.  The student / teacher ``Parameter`` objects stay the
ones the model owns: their ``.data`` is re-pointed at views of the flat buffers (``state_dict()``,
checkpoints and ``OldWeightEMA`` keep working), and ``p.grad`` is a view of the flat gradient bucket, so
backward writes straight into peer-readable memory.  ``PeerGroup.virtual`` builds W ranks inside ONE process
on one GPU (same kernels, same protocol; the "peers" are local buffers) — that is how the N-rank arithmetic
is parity-tested on a single-GPU box.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

__all__ = ["PeerGroup", "ShardedStudentStep", "flat_layout", "READY", "REDUCED", "PARAMS", "COUNTS"]

READY, REDUCED, PARAMS, COUNTS = 0, 1, 2, 3
PAD_BYTES = 8192
PAD_ERR = 4 * 8 + 8          # UDAPE_DP_PAD_ERR
MAX_RANKS = 8
_ALIGN = 256


class _DeviceSpan:
    """``__cuda_array_interface__`` holder: lets torch view memory this package cudaMalloc'ed or IPC-mapped."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.owner = owner   # keeps the arena alive as long as any tensor view exists
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def _round_up(n: int, a: int) -> int:
    return (n + a - 1) // a * a


def flat_layout(params) -> tuple[list[int], int]:
    """Element offsets of every parameter inside a flat float32 buffer (each tensor starts on a 16-byte
    boundary) and the padded total (a multiple of 4 elements)."""
    offsets, total = [], 0
    for p in params:
        total = _round_up(total, 4)
        offsets.append(total)
        total += p.numel()
    return offsets, _round_up(total, 4)


def arena_bytes(n_total: int, world: int = 1) -> int:
    """pad | gradient bucket | student parameters | shadow of the rank's slice, each 256-byte aligned.  (The
    shadow is sized for world = 1, the largest slice, so one arena serves any world size.)"""
    return PAD_BYTES + 3 * _round_up(4 * _round_up(n_total, 4096), _ALIGN)


class PeerGroup:
    """This rank's peer-mappable arena plus the addresses of every other rank's arena as mapped here."""

    def __init__(self, rank: int, world: int, dev: torch.device, nbytes: int, bases: list[int], owned: list[int],
                 opened: list[int]):
        self.rank, self.world, self.dev, self.nbytes = rank, world, dev, nbytes
        self.bases = bases            # bases[q]: arena of rank q in THIS process's address space
        self._owned, self._opened = owned, opened   # to cudaFree / to cudaIpcCloseMemHandle on close()
        self._closed = False

    # -- construction ------------------------------------------------------------------------------
    @staticmethod
    def _alloc(nbytes: int, dev: torch.device) -> int:
        p = ctypes.c_void_p()
        with _lib.on_device(dev):
            _lib.check(_lib.load().udape_peer_alloc(nbytes, ctypes.byref(p)), "udape_peer_alloc")
        return int(p.value)

    @classmethod
    def create(cls, nbytes: int, dev: torch.device, group=None) -> "PeerGroup":
        """Collective over ``group`` (default: the world): allocate, exchange IPC handles, map the peers."""
        import torch.distributed as dist

        dev = torch.device(dev)
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            base = cls._alloc(nbytes, dev)
            return cls(0, 1, dev, nbytes, [base], [base], [])
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world > MAX_RANKS:
            raise ValueError(f"PeerGroup: at most {MAX_RANKS} ranks (one NVSwitch domain), got {world}")
        lib = _lib.load()
        # every rank reaches both exchanges whatever happens locally, and all ranks fail together: a rank that
        # raised alone would leave the others waiting in the next collective
        base, handle, err = None, (ctypes.c_ubyte * 64)(), None
        try:
            base = cls._alloc(nbytes, dev)
            with _lib.on_device(dev):
                _lib.check(lib.udape_peer_export(base, handle), "udape_peer_export")
        except Exception as exc:
            err = f"rank {rank}: {type(exc).__name__}: {exc}"
        everyone = [None] * world
        dist.all_gather_object(everyone, (bytes(handle), int(nbytes), err), group=group)
        bases, opened = [], []
        if not any(e for _, _, e in everyone):
            try:
                for q, (h, nb, _) in enumerate(everyone):
                    if nb != nbytes:
                        raise RuntimeError(f"rank {q} allocated {nb} bytes, this rank {nbytes} (layouts must match)")
                    if q == rank:
                        bases.append(base)
                        continue
                    p = ctypes.c_void_p()
                    with _lib.on_device(dev):
                        _lib.check(lib.udape_peer_open((ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(p)),
                                   f"udape_peer_open(rank {q})")
                    bases.append(int(p.value))
                    opened.append(int(p.value))
            except Exception as exc:
                err = f"rank {rank}: {type(exc).__name__}: {exc}"
        errors = [None] * world
        dist.all_gather_object(errors, err, group=group)   # also the barrier: every rank has mapped every arena
        failed = [e for e in errors if e] + [e for _, _, e in everyone if e]
        if failed:
            with _lib.on_device(dev):
                for p in opened:
                    lib.udape_peer_close(p)
                if base is not None:
                    lib.udape_peer_free(base)
            raise _lib.UdapeError("PeerGroup.create: peer memory could not be set up: " + "; ".join(sorted(set(failed))))
        return cls(rank, world, dev, nbytes, bases, [base], opened)

    @classmethod
    def virtual(cls, world: int, nbytes: int, dev: torch.device) -> list["PeerGroup"]:
        """``world`` ranks inside this process on one device (tests, single-GPU validation of the protocol)."""
        dev = torch.device(dev)
        bases = [cls._alloc(nbytes, dev) for _ in range(world)]
        return [cls(r, world, dev, nbytes, bases, bases if r == 0 else [], []) for r in range(world)]

    # -- views -------------------------------------------------------------------------------------
    def tensor(self, q: int, offset: int, numel: int, dtype=torch.float32) -> torch.Tensor:
        """1-D view of ``numel`` elements at byte ``offset`` of rank ``q``'s arena."""
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        if offset < 0 or offset + nbytes > self.nbytes:
            raise ValueError("PeerGroup.tensor: span outside the arena")
        with torch.cuda.device(self.dev):
            raw = torch.as_tensor(_DeviceSpan(self.bases[q] + offset, nbytes, self), device=self.dev)
        return raw.view(dtype)

    def error_word(self) -> int:
        """Non-zero after a bounded wait expired (1 + phase).  Host sync."""
        return int(self.tensor(self.rank, 4 * PAD_ERR, 1, torch.int32).item())

    def close(self):
        if self._closed:
            return
        self._closed = True
        lib = _lib.load()
        with _lib.on_device(self.dev):
            torch.cuda.synchronize(self.dev)
            for p in self._opened:
                lib.udape_peer_close(p)
            for p in self._owned:      # virtual ranks: rank 0 owns every arena
                lib.udape_peer_free(p)


class ShardedStudentStep(torch.optim.Optimizer):
    """``stu_optimizer`` + ``tea_optimizer`` of ``train_human.py:136-141`` for one-process-per-GPU training:
    ``torch.optim.Adam(lr, betas, eps, weight_decay)`` or ``SGD(lr, momentum, dampening, weight_decay,
    nesterov)`` on the rank-averaged gradient, with ``OldWeightEMA(alpha)`` folded into the parameter
    exchange.  One param group (what the trainers build).  ``step()`` is four launches, no host sync, CUDA-graph
    capturable; ``torch.amp.GradScaler`` support through ``uda_poseestimation_b200.GradScaler``
    (``found_inf`` is produced inside the step, on the reduced gradient).

    ``peers`` decides the world: ``PeerGroup.create`` (one process per GPU) or one element of
    ``PeerGroup.virtual``.  ``teacher_params=None`` runs without the EMA (the reference's ``pretrain()``)."""

    _step_supports_amp_scaling = True

    def __init__(self, params, peers: PeerGroup, algo: str = "adam", teacher_params=None, alpha: float = 0.999,
                 lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 momentum: float = 0.0, dampening: float = 0.0, nesterov: bool = False, timeout_s: float = 20.0,
                 capturable: bool = False):
        if algo not in ("adam", "sgd"):
            raise ValueError("algo must be 'adam' or 'sgd'")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, momentum=momentum,
                        dampening=dampening, nesterov=nesterov)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("ShardedStudentStep supports one param group (train_human.py:136-139 builds one)")
        self.algo, self.peers, self.alpha, self.capturable = algo, peers, alpha, capturable
        self.timeout_ns = int(timeout_s * 1e9)
        ps = self.param_groups[0]["params"]
        dev = peers.dev
        for p in ps:
            if p.device != dev or p.dtype != torch.float32:
                raise TypeError("ShardedStudentStep: parameters must be float32 on the peer group's device")
        self.offsets, self.n_total = flat_layout(ps)
        if arena_bytes(self.n_total) > peers.nbytes:
            raise ValueError(f"peer arena too small: {peers.nbytes} < {arena_bytes(self.n_total)} bytes")
        lib = _lib.load()
        self.shard_elems = int(lib.udape_dp_shard_elems(self.n_total, peers.world))
        self._g_off = PAD_BYTES
        self._p_off = self._g_off + _round_up(4 * self.n_total, _ALIGN)
        self._s_off = self._p_off + _round_up(4 * self.n_total, _ALIGN)
        if self._s_off + 4 * self.shard_elems > peers.nbytes:
            raise ValueError(f"peer arena too small for the shadow slice: {peers.nbytes} bytes")
        self.flat_grads = peers.tensor(peers.rank, self._g_off, self.n_total)
        self.flat_params = peers.tensor(peers.rank, self._p_off, self.n_total)
        with torch.no_grad():
            for p, off in zip(ps, self.offsets):
                view = self.flat_params[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view                                     # the model's Parameter objects, new storage
                p.grad = self.flat_grads[off:off + p.numel()].view(p.shape)
        self.flat_teacher = None
        if teacher_params is not None:
            tps = list(teacher_params)
            if len(tps) != len(ps) or any(t.shape != p.shape or t.dtype != p.dtype or t.device != dev for t, p in zip(tps, ps)):
                raise ValueError("teacher / student parameter lists do not match")
            self.flat_teacher = torch.zeros(self.n_total, dtype=torch.float32, device=dev)
            with torch.no_grad():
                for t, off in zip(tps, self.offsets):
                    view = self.flat_teacher[off:off + t.numel()].view(t.shape)
                    view.copy_(t.data)
                    t.data = view
        self.shadow = peers.tensor(peers.rank, self._s_off, self.shard_elems)
        need1 = algo == "adam" or momentum != 0
        # [2, S]: the update writes the half that is not current; committing a step flips them (step count parity)
        self.state1 = torch.zeros((2, self.shard_elems), dtype=torch.float32, device=dev) if need1 else None
        self.state2 = torch.zeros((2, self.shard_elems), dtype=torch.float32, device=dev) if algo == "adam" else None
        self.found_inf = torch.zeros((), dtype=torch.float32, device=dev)
        self._step_dev = torch.zeros((), dtype=torch.int32, device=dev)
        self._words = torch.zeros(8, dtype=torch.int32, device=dev)   # epoch, counts epoch, ws[2], ticket K2, ticket K3
        self._lr_dev = torch.empty((), dtype=torch.float32, device=dev) if capturable else None
        self._c = _lib.DpPeers()
        self._c.rank, self._c.world = peers.rank, peers.world
        for q in range(peers.world):
            self._c.pads[q] = peers.bases[q]
            self._c.grads[q] = peers.bases[q] + self._g_off
            self._c.params[q] = peers.bases[q] + self._p_off
            self._c.shadow[q] = peers.bases[q] + self._s_off
        if capturable:
            self.sync_lr()

    # -- helpers -----------------------------------------------------------------------------------
    def _word(self, i: int) -> int:
        return self._words.data_ptr() + 4 * i

    def sync_lr(self):
        self._lr_dev.fill_(float(self.param_groups[0]["lr"]))

    def shard_bounds(self, rank: int | None = None) -> tuple[int, int]:
        r = self.peers.rank if rank is None else rank
        lo = min(r * self.shard_elems, self.n_total)
        return lo, min(lo + self.shard_elems, self.n_total)

    def _hyper(self) -> "_lib.OptHyper":
        g = self.param_groups[0]
        h = _lib.OptHyper()
        h.lr, h.eps, h.weight_decay, h.step = float(g["lr"]), float(g["eps"]), float(g["weight_decay"]), 1
        if self.algo == "adam":
            h.beta1, h.beta2 = g["betas"]
            h.nesterov = 0
        else:
            h.beta1, h.beta2, h.nesterov = float(g["momentum"]), float(g["dampening"]), int(bool(g["nesterov"]))
        h.ema_a, h.ema_b = float(self.alpha), float(1.0 - self.alpha)      # utils.py:22
        return h

    # -- the step ------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, _events=None):
        """``_events`` (profiling, eager only): five CUDA events recorded around the four launches."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ev = iter(_events) if _events is not None else None
        self.step_begin(ev)
        self.step_finish(None, ev)
        return loss

    def _launch_args(self):
        dev = self.peers.dev
        grad_scale = getattr(self, "grad_scale", None)
        if grad_scale is not None and (grad_scale.device != dev or grad_scale.dtype != torch.float32):
            raise RuntimeError("grad_scale must be a float32 tensor on the parameters' device")
        return _lib.load(), dev, grad_scale, ctypes.byref(self._c), self._word(0), self.timeout_ns

    @staticmethod
    def _mark(ev, dev):
        if ev is not None:
            next(ev).record(torch.cuda.current_stream(dev))

    @torch.no_grad()
    def step_begin(self, _ev=None):
        """First half of ``step()``: the READY barrier and the gradient reduce-scatter + update.  A caller that has other
        work to enqueue (e.g. the PCK launch whose counts ``step_finish`` exchanges) does it between the halves."""
        lib, dev, grad_scale, c, epoch, t_ns = self._launch_args()
        with _lib.on_device(dev):
            st = _lib.stream_ptr(dev)
            self._mark(_ev, dev)
            _lib.check(lib.udape_dp_barrier(c, READY, epoch, t_ns, st), "udape_dp_barrier")
            self._mark(_ev, dev)
            h = self._hyper()
            algo = _lib.OPT_ADAM if self.algo == "adam" else _lib.OPT_SGD
            _lib.check(lib.udape_dp_reduce_step(c, self.n_total, algo, ctypes.byref(h), _lib.ptr(self._lr_dev),
                                                _lib.ptr(grad_scale), self._step_dev.data_ptr(), _lib.ptr(self.state1),
                                                _lib.ptr(self.state2), epoch, self._word(2), st), "udape_dp_reduce_step")
            self._mark(_ev, dev)

    @torch.no_grad()
    def step_finish(self, counts: torch.Tensor | None = None, _ev=None):
        """Second half: [the PCK-count exchange of ``counts`` (in place), on this same stream so that every rank
        meets its peers in ONE order and no two waiting kernels of a rank can ever wait for each other,] the
        non-finite verdict and the parameter all-gather (the commit) + EMA."""
        lib, dev, _, c, epoch, t_ns = self._launch_args()
        h = self._hyper()
        if counts is not None:
            self.allreduce_counts(counts, out=counts)
        with _lib.on_device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(lib.udape_dp_wait(c, REDUCED, epoch, self.found_inf.data_ptr(), t_ns, st), "udape_dp_wait")
            self._mark(_ev, dev)
            _lib.check(lib.udape_dp_gather_ema(c, self.n_total, _lib.ptr(self.flat_teacher), h.ema_a, h.ema_b,
                                               self.found_inf.data_ptr(), self._step_dev.data_ptr(), epoch, self._word(5),
                                               st), "udape_dp_gather_ema")
            self._mark(_ev, dev)

    def state_shards(self):
        """This rank's CURRENT optimizer-state shards ``(exp_avg | momentum_buffer, exp_avg_sq)`` (views, host sync)."""
        cur = self.applied_steps() & 1
        return (self.state1[cur] if self.state1 is not None else None, self.state2[cur] if self.state2 is not None else None)

    kernels_per_step = 4

    def allreduce_counts(self, counts: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """SUM over ranks of an int32 tensor of <= 64 elements (PCK ``hits || valid``) in one single-CTA launch
        over the signal pads; returns ``out`` (a new tensor by default; in place when ``out is counts``)."""
        if counts.dtype != torch.int32 or counts.numel() > 64 or not counts.is_contiguous():
            raise TypeError("allreduce_counts: contiguous int32 tensor of at most 64 elements")
        dev = _lib.require_cuda(counts)
        out = torch.empty_like(counts) if out is None else out
        with _lib.on_device(dev):
            _lib.check(_lib.load().udape_dp_allreduce_counts(ctypes.byref(self._c), counts.data_ptr(), counts.numel(),
                                                             out.data_ptr(), self._word(1), self.timeout_ns,
                                                             _lib.stream_ptr(dev)), "udape_dp_allreduce_counts")
        return out

    def zero_grad(self, set_to_none: bool = False):
        """Zeroes the flat bucket in place (the gradients ARE the peer-readable bucket: never freed)."""
        self.flat_grads.zero_()

    def applied_steps(self) -> int:
        return int(self._step_dev.item())

    def check(self):
        """Raise if a bounded cross-rank wait expired since the last call (host sync)."""
        code = self.peers.error_word()
        if code:
            self.peers.tensor(self.peers.rank, 4 * PAD_ERR, 1, torch.int32).zero_()
            names = {1: "READY", 2: "REDUCED", 3: "PARAMS", 4: "COUNTS"}
            raise _lib.UdapeError(f"data-parallel step: wait for phase {names.get(code, code)} timed out on rank "
                                  f"{self.peers.rank} (a rank is missing or out of step)")

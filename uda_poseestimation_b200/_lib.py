"""ctypes binding of ``libudape_b200.so`` (the C-ABI declared in ``include/udape.h``).

There is no CPU fallback and no alternative backend: if the shared library is missing it
is rebuilt with nvcc when possible, otherwise importing any operator raises.  Every
wrapper passes raw ``data_ptr()``s plus the caller's *current* CUDA stream; ctypes drops
the GIL for the duration of the call, so ``nn.DataParallel``-style multi-threaded callers
are safe (the library keeps no global state besides a thread-local error string).
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

import torch

from . import build as _build

# dtype codes of include/udape.h
F32, F16, BF16, U8 = 0, 1, 2, 3
_DTYPE_CODE = {
    torch.float32: F32,
    torch.float16: F16,
    torch.bfloat16: BF16,
    torch.uint8: U8,
    torch.bool: U8,
}

ERR_NAMES = {-1: "NULL pointer", -2: "unsupported dtype", -3: "bad shape", -4: "misaligned pointer",
             -5: "argument out of range"}


class UdapeError(RuntimeError):
    """Raised for any non-zero status returned by the C-ABI."""


class EmaChunk(ctypes.Structure):
    _fields_ = [("dst", c_void_p), ("src", c_void_p), ("numel", c_int64)]


class OptChunk(ctypes.Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("state1", c_void_p), ("state2", c_void_p),
                ("ema", c_void_p), ("fresh", c_void_p), ("numel", c_int64)]


class OptHyper(ctypes.Structure):
    _fields_ = [("lr", c_double), ("beta1", c_double), ("beta2", c_double), ("eps", c_double),
                ("weight_decay", c_double), ("ema_a", c_float), ("ema_b", c_float),
                ("step", ctypes.c_int32), ("nesterov", ctypes.c_int32)]


class AdainJob(ctypes.Structure):
    _fields_ = [("content", c_void_p), ("style", c_void_p), ("out", c_void_p), ("alpha_dev", c_void_p), ("alpha", c_float)]


class TargetJob(ctypes.Structure):
    _fields_ = [("joints", c_void_p), ("vis", c_void_p), ("target", c_void_p), ("weight", c_void_p),
                ("hm_w", ctypes.c_int32), ("hm_h", ctypes.c_int32)]


class LabelmapJob(ctypes.Structure):
    _fields_ = [("pts", c_void_p), ("gate", c_void_p), ("img", c_void_p), ("vis_out", c_void_p)]


class DpPeers(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("grads", c_void_p * 8), ("params", c_void_p * 8),
                ("shadow", c_void_p * 8), ("pads", c_void_p * 8)]


OPT_ADAM, OPT_SGD = 0, 1


# name -> (restype, argtypes); must list every UDAPE_API symbol of include/udape.h
PROTOTYPES = {
    "udape_version": (c_int, []),
    "udape_build_info": (c_char_p, []),
    "udape_last_error": (c_int, [c_char_p, c_size_t]),
    "udape_table_feed": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "udape_mean_std": (c_int, [c_void_p, c_int, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "udape_mean_std_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p,
                                   c_void_p]),
    "udape_adain_mix": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_float, c_float,
                                c_void_p, c_void_p, c_void_p]),
    "udape_adain_mix_multi": (c_int, [POINTER(AdainJob), c_int, c_int, c_int64, c_int64, c_int64, c_float, c_void_p]),
    "udape_channel_clamp": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "udape_decode": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_float, c_void_p, c_double, c_void_p, c_void_p]),
    "udape_decode_select": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_float, c_void_p, c_double, c_void_p, c_int64, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "udape_mask_select": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "udape_pck_counts": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64,
                                 c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "udape_joints_mse_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_int64,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "udape_joints_mse_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_int64,
                                     c_void_p, c_int, c_void_p, c_void_p]),
    "udape_cons_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64,
                               c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "udape_cons_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64,
                               c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "udape_loss_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                                c_double, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int, c_int,
                                c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "udape_gauss_target": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_double, c_double,
                                   c_double, c_void_p, c_void_p, c_void_p]),
    "udape_labelmap": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_double, c_int, c_int, c_void_p,
                               c_void_p, c_void_p]),
    "udape_gauss_target_multi": (c_int, [POINTER(TargetJob), c_int, c_int64, c_double, c_double, c_double, c_void_p]),
    "udape_labelmap_multi": (c_int, [POINTER(LabelmapJob), c_int, c_int64, c_int64, c_int64, c_double, c_int, c_void_p]),
    "udape_ema_plan": (c_int64, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int64), c_int64, c_int64,
                                 c_int64, POINTER(EmaChunk), c_int64]),
    "udape_ema_multi": (c_int, [c_void_p, c_int64, c_int64, c_float, c_float, c_int, c_int, c_void_p]),
    "udape_opt_plan": (c_int64, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                 POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int64), c_int64, c_int64,
                                 POINTER(OptChunk), c_int64]),
    "udape_grad_check": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "udape_student_step": (c_int, [c_void_p, c_int64, c_int, POINTER(OptHyper), c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "udape_dp_shard_elems": (c_int64, [c_int64, c_int]),
    "udape_dp_barrier": (c_int, [POINTER(DpPeers), c_int, c_void_p, ctypes.c_uint64, c_void_p]),
    "udape_dp_wait": (c_int, [POINTER(DpPeers), c_int, c_void_p, c_void_p, ctypes.c_uint64, c_void_p]),
    "udape_dp_reduce_step": (c_int, [POINTER(DpPeers), c_int64, c_int, POINTER(OptHyper), c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "udape_dp_gather_ema": (c_int, [POINTER(DpPeers), c_int64, c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "udape_dp_allreduce_counts": (c_int, [POINTER(DpPeers), c_void_p, c_int, c_void_p, c_void_p, ctypes.c_uint64,
                                          c_void_p]),
    "udape_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "udape_peer_free": (c_int, [c_void_p]),
    "udape_peer_export": (c_int, [c_void_p, POINTER(ctypes.c_ubyte)]),
    "udape_peer_open": (c_int, [POINTER(ctypes.c_ubyte), POINTER(c_void_p)]),
    "udape_peer_close": (c_int, [c_void_p]),
    "udape_rewarp_fwd": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_int, c_int, c_int, c_int, c_void_p, c_int,
                                 c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "udape_rewarp_plan_elems": (c_int64, [c_int64, c_int64, c_int]),
    "udape_rewarp_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_int,
                                 c_void_p, c_void_p, c_void_p]),
    "udape_rewarp_decode_select": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_int,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int64, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lock = threading.Lock()
_lib = None


def library_path() -> str:
    return str(_build.lib_path())


def load(rebuild: bool = True) -> ctypes.CDLL:
    """Load (building first if stale and nvcc is available) and type the shared library."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.lib_path()
        if rebuild and os.environ.get("UDAPE_NO_BUILD") != "1":
            try:
                if _build.needs_build():
                    _build.build()
            except Exception as exc:  # no nvcc on this box: use the shipped .so if any
                if not path.exists():
                    raise ImportError(
                        f"libudape_b200.so is missing and could not be built: {exc}\n"
                        "There is no CPU/PyTorch fallback; run `python -m uda_poseestimation_b200.build`."
                    ) from exc
        if not path.exists():
            raise ImportError(f"{path} not found; run `python -m uda_poseestimation_b200.build`")
        lib = ctypes.CDLL(str(path))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export the symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    load().udape_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status == 0:
        return
    msg = last_error()
    if status < 0:
        raise ValueError(f"{what}: {ERR_NAMES.get(status, status)}: {msg}")
    raise UdapeError(f"{what}: CUDA error {status}: {msg}")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPE_CODE[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype} (expected float32/float16/bfloat16)") from None


def float_code(t: torch.Tensor) -> int:
    c = dtype_code(t)
    if c == U8:
        raise TypeError(f"unsupported dtype {t.dtype} (expected float32/float16/bfloat16)")
    return c


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """All tensors must live on the same CUDA device (there is no CPU path)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "uda_poseestimation_b200 operators are CUDA-only (sm_100a); got a tensor on "
                f"{t.device}. There is no CPU fallback — use the reference implementation for CPU tensors."
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class on_device:
    """Make ``dev`` the current CUDA device for the duration of a launch (cheap no-op when
    it already is, which is the common case)."""

    __slots__ = ("idx", "prev")

    def __init__(self, dev: torch.device):
        self.idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self.prev = -1

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def no_autograd(name: str, *tensors: torch.Tensor) -> None:
    """Forward-only operators fail loudly instead of silently detaching the graph."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            f"{name}: the CUDA operator is forward-only (the trainers call it under torch.no_grad(), "
            "train_human.py:347); wrap the call in torch.no_grad() or detach the inputs"
        )


class _TicketPool:
    """Zero-initialised uint32 words for the kernels' last-CTA-done tickets (include/udape.h:
    "ticket must be zero on entry; the call leaves it zero").  The words are self-resetting, so a
    slot can be handed out again without a memset; slots rotate so that kernels running
    concurrently on different streams never share one, and a slot baked into a CUDA graph during
    stream capture is retired from the rotation for good."""

    SLOTS = 4096

    def __init__(self, dev: torch.device):
        self.buf = torch.zeros(self.SLOTS, dtype=torch.int32, device=dev)
        self.base = self.buf.data_ptr()
        self.reserved = 0          # [0, reserved) belong to captured graphs
        self.cursor = 0            # rotation over [reserved, SLOTS)
        self.spill = []            # extra buffers once graphs have eaten half the pool

    def take(self) -> int:
        if torch.cuda.is_current_stream_capturing():
            if self.reserved >= self.SLOTS // 2:
                extra = torch.zeros(1, dtype=torch.int32, device=self.buf.device)
                self.spill.append(extra)  # kept alive for the life of the process, like the graph
                return extra.data_ptr()
            slot = self.reserved
            self.reserved += 1
            return self.base + 4 * slot
        span = self.SLOTS - self.reserved
        slot = self.reserved + self.cursor % span
        self.cursor += 1
        return self.base + 4 * slot


_ticket_pools: dict = {}


def check_tickets(reset: bool = False) -> None:
    """Every ticket word must be zero between launches ("ticket must be zero on entry; the call leaves it zero").
    A launch that aborted mid-grid (a sticky CUDA error, a killed context) can leave one non-zero, and every later
    reduction that draws that slot would elect the wrong "last CTA" without any other symptom.  Call this after
    recovering from a CUDA error, or at checkpoints (it synchronises the device): raises ``UdapeError`` naming the
    dirty slots; ``reset=True`` zeroes them instead (only when no launch is in flight)."""
    with _lock:
        pools = list(_ticket_pools.items())
    for idx, pool in pools:
        with torch.cuda.device(idx):
            torch.cuda.synchronize(idx)
            bufs = [pool.buf] + list(pool.spill)
            dirty = [(i, int(n)) for i, b in enumerate(bufs) for n in torch.nonzero(b).flatten().tolist()]
            if dirty and reset:
                for b in bufs:
                    b.zero_()
            elif dirty:
                raise UdapeError(f"cuda:{idx}: {len(dirty)} ticket word(s) are not zero between launches (buffer, slot): "
                                 f"{dirty[:8]} — a launch aborted mid-grid; call check_tickets(reset=True) after recovery")


def ticket(dev: torch.device) -> int:
    """Device address of a zeroed, self-resetting uint32 ticket word on ``dev``."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    with _lock:
        pool = _ticket_pools.get(idx)
        if pool is None:
            with torch.cuda.device(idx):
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("the first loss/PCK call on a device must happen before CUDA-graph capture "
                                       "(it allocates the ticket pool); run one warm-up step first")
                pool = _ticket_pools[idx] = _TicketPool(torch.device("cuda", idx))
        return pool.take()

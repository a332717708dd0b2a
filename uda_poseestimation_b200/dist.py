"""Multi-GPU plumbing: one process per GPU, batch sharded along dim 0, no data-path
collective inside any hot-path kernel (every plane is independent).

The reference is single-process ``nn.DataParallel`` (``train_human.py:145-148``): it
re-broadcasts parameters every forward, gathers outputs on GPU 0 and runs every loss, mask,
EMA and PCK there.  Here each rank runs the whole hot path on its own shard and NCCL over
NVLink is used for exactly two exchanges:

* the integer PCK-count all-reduce — ``hits[K] ‖ valid[K]`` int32, summed before the per-joint
  ratios are formed (``accuracy`` is a mean of per-joint ratios, keypoint_detection.py:86-92,
  so ratios cannot be averaged across ranks);
* the gradient all-reduce of the student parameters (one flat bucket, so a single NCCL call).

``torch.distributed`` is only the transport; the ``gloo`` backend with CPU tensors is
supported for the host-logic tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .keypoint_detection import _pck, accuracy_from_counts

__all__ = ["init_from_env", "bind_to_gpu_numa", "shard_bounds", "shard", "allreduce_counts", "distributed_accuracy",
           "FlatGradBucket", "mean_scalar"]


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise the default process group from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*.
    Returns ``(rank, world_size, local_rank)``; a no-op for single-process runs."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def bind_to_gpu_numa(local_rank: int) -> list[int] | None:
    """Pin this process to the CPUs NVML reports as local to GPU ``local_rank`` (its NUMA node / PCIe
    root).  Call before allocating pinned host buffers: first touch then places them in the memory of
    that node, so with one process per GPU the ranks' host->device copies do not all pull from one
    socket (8 ranks x 52 GB/s exceeds a single socket's memory and inter-socket bandwidth).  Returns the
    CPU list, or None when NVML / affinity is unavailable (the call is best-effort and never raises)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = local_rank
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                index = int(ids[local_rank])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = (n_cpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of ``n`` samples: the first ``n % world`` ranks get one extra."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's slice of a global batch tensor (dim 0)."""
    s, e = shard_bounds(t.shape[0], rank, world)
    return t[s:e]


def allreduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of an integer count tensor (``hits ‖ valid``, int32)."""
    if counts.dtype not in (torch.int32, torch.int64):
        raise TypeError("allreduce_counts expects an integer tensor")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def distributed_accuracy(output: torch.Tensor, target: torch.Tensor, thr: float = 0.5, group=None):
    """PCK over the GLOBAL batch from per-rank shards: local integer counts → one int32
    all-reduce → the reference's ``(acc[K], avg_acc, cnt)`` (identical on every rank, and equal
    to ``accuracy`` on the concatenated batch), plus this rank's ``pred``."""
    counts, pred = _pck(output, target, thr)
    allreduce_counts(counts, group)
    host = counts.cpu().numpy()
    acc, avg_acc, cnt = accuracy_from_counts(host[0], host[1])
    return acc, avg_acc, cnt, pred


def mean_scalar(x: torch.Tensor, group=None) -> torch.Tensor:
    """Average of a per-rank scalar (equal shard sizes ⇒ the global-batch mean loss)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        x = x.clone()
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        x /= dist.get_world_size(group)
    return x


class FlatGradBucket:
    """All student gradients in ONE contiguous buffer: ``p.grad`` of every parameter is a view
    into it, so backward writes straight into the bucket and the data-parallel reduction is a
    single NCCL all-reduce (212 MB fp32 for PoseResNet-101) with no gather/scatter copies."""

    def __init__(self, params, dtype: torch.dtype | None = None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket: no parameters require grad")
        dev = self.params[0].device
        dtype = dtype or self.params[0].dtype
        offsets, total = [], 0
        for p in self.params:
            if p.device != dev:
                raise ValueError("FlatGradBucket: parameters on different devices")
            total = (total + 3) // 4 * 4  # keep every view 16-byte aligned for fp32
            offsets.append(total)
            total += p.numel()
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        self.views = []
        for p, off in zip(self.params, offsets):
            v = self.flat[off:off + p.numel()].view(p.shape)
            p.grad = v
            self.views.append(v)

    def zero_(self):
        self.flat.zero_()

    def allreduce_(self, average: bool = True, group=None, async_op: bool = False):
        """SUM (or mean) of the bucket across ranks, in place. Returns the work handle if async."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        if average:
            # pre-scale so the collective itself stays a pure SUM (works on gloo and nccl alike)
            self.flat.div_(dist.get_world_size(group))
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return work if async_op else None

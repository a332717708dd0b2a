"""How fast can one GPU's SMs pull another GPU's memory over NVLink?  128-bit loads vs cp.async.bulk (1-D TMA),
every rank pulling from its neighbour at the same time (what the data-parallel tail does), beside a copy-engine
copy of the same bytes.  torchrun --nproc-per-node N tools/probe/peer_probe.py [--out file.json]"""
import argparse
import ctypes
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from uda_poseestimation_b200 import _lib  # noqa: E402
from uda_poseestimation_b200.dp import PeerGroup  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--mb", type=int, default=192)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    nbytes = args.mb << 20
    peers = PeerGroup.create(nbytes, dev)
    peers.tensor(rank, 0, nbytes // 4).normal_()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    lib = ctypes.CDLL(str(ROOT / "build" / "peer_probe.so"))
    lib.probe_pull.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.probe_pull.restype = ctypes.c_int
    lib.probe_push.argtypes = lib.probe_pull.argtypes
    lib.probe_push.restype = ctypes.c_int
    torch.cuda.synchronize()
    dist.barrier()
    rows = []

    def timed(name, fn, reps=10, scale=1.0):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        gbs = scale * nbytes / (ms.item() * 1e-3) / 1e9
        rows.append({"what": name, "ms": ms.item(), "GBps_per_gpu": gbs})
        if rank == 0:
            print(f"{name:60s} {ms.item():8.3f} ms  {gbs:8.1f} GB/s per GPU", flush=True)

    st = lambda: _lib.stream_ptr(dev)
    for who, q in (("local", rank), ("peer", (rank + 1) % world)):
        src = peers.bases[q]
        if who == "peer" and world == 1:
            continue
        timed(f"{who}: copy engine / torch copy_", lambda: dst.view(torch.float32).copy_(peers.tensor(q, 0, nbytes // 4)))
        for un in (4,):
            for grid in (148 * 4,):
                timed(f"{who}: 128-bit loads, {un} per thread, grid {grid}", lambda: lib.probe_pull(src, dst.data_ptr(), nbytes, 0, grid, un, st()))
        for chunk in (16384,):
            for grid in (148 * 2,):
                if 4 * chunk * (grid // 148) > 200 * 1024:
                    continue
                timed(f"{who}: cp.async.bulk, {chunk} B chunks x 4 stages, grid {grid}", lambda: lib.probe_pull(src, dst.data_ptr(), nbytes, 1, grid, chunk, st()))
        # correctness of the last TMA pull
        torch.cuda.synchronize()
        dist.barrier()
        ok = torch.equal(dst.view(torch.float32), peers.tensor(q, 0, nbytes // 4))
        if rank == 0:
            print(f"{who}: last pull equals the source: {ok}")
        # push: local source, the neighbour's arena (second half) as the destination
        if who == "peer":
            half = nbytes // 2
            lsrc, pdst = peers.bases[rank], peers.bases[q] + half
            for un in (4, 8):
                for grid in (148 * 2, 148 * 8):
                    timed(f"push: 128-bit stores, {un} per thread, grid {grid} ({half >> 20} MB)", lambda: lib.probe_push(lsrc, pdst, half, 0, grid, un, st()), scale=0.5)
            for chunk in (16384, 32768):
                for grid in (148 * 2, 148 * 4):
                    timed(f"push: cp.async.bulk load + bulk store, {chunk} B, grid {grid} ({half >> 20} MB)", lambda: lib.probe_push(lsrc, pdst, half, 1, grid, chunk, st()), scale=0.5)
            timed(f"push: copy engine ({half >> 20} MB)", lambda: peers.tensor(q, half, half // 4).copy_(peers.tensor(rank, 0, half // 4)), scale=0.5)
            torch.cuda.synchronize()
            dist.barrier()
            okp = torch.equal(peers.tensor(rank, half, half // 4), peers.tensor((rank - 1) % world, 0, half // 4))
            if rank == 0:
                print(f"push: what the neighbour pushed here equals its source: {okp}")
    if rank == 0 and args.out:
        json.dump({"world": world, "mbytes": args.mb, "rows": rows}, open(args.out, "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

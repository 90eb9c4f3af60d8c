"""Multi-GPU plumbing for the bootstrap: replicate ids are split contiguously across ranks (one
process per GPU), every rank holds the full observation matrix, and ONE all-gather of the
per-replicate result rows ends the run (reference: fork + Queue, bootstrap.py:91-113).
torch.distributed is plumbing only: NCCL on CUDA tensors, gloo on CPU tensors (tests)."""
from __future__ import annotations

import os
import sys

import numpy as np


def _dist():
    if "torch" not in sys.modules and int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return None
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def rank_world():
    dist = _dist()
    return (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)


def shard_range(total: int, rank: int, world: int):
    """Contiguous [begin, begin+count) of global replicate ids for `rank`; the first total % world
    ranks take one extra replicate."""
    base, extra = divmod(int(total), int(world))
    count = base + (1 if rank < extra else 0)
    begin = rank * base + min(rank, extra)
    return begin, count


def broadcast_int(value: int, src: int = 0) -> int:
    dist = _dist()
    if not dist:
        return int(value)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    dist.broadcast(t, src)
    return int(t.item())


def send_buffer(total: int, width: int):
    """Flat float64 CUDA tensor [cap*width rows | cap status | cap iters] the engine writes its rows
    into directly (NCCL backend), or None when not distributed / gloo.  cap = largest shard, so that
    ONE all_gather_into_tensor moves everything."""
    dist = _dist()
    if not dist or dist.get_backend() != "nccl":
        return None
    import torch
    cap = shard_range(total, 0, dist.get_world_size())[1]
    buf = torch.zeros(cap * (width + 2), dtype=torch.float64, device="cuda")
    # the engine writes rows into this buffer on its own (non-blocking) stream: the zero fill on torch's
    # stream has to be complete before the pointer is handed over
    torch.cuda.current_stream().synchronize()
    return buf


def allgather_rows(rows, status, iters, total: int, width: int, device_buffer=None):
    """Gathers (rows [count, width], status, iters) of all ranks in global replicate order with one
    collective.  rows may be None when `device_buffer` (from send_buffer) already holds them."""
    dist = _dist()
    if not dist:
        return rows, status, iters
    import torch
    world = dist.get_world_size()
    cap = shard_range(total, 0, world)[1]
    count = len(status)
    tail = np.zeros(2 * cap, dtype=np.float64)
    tail[:count] = status
    tail[cap:cap + count] = iters
    if device_buffer is not None:
        send = device_buffer
        send[cap * width:] = torch.from_numpy(tail).to(send.device)
    else:
        flat = np.zeros(cap * (width + 2), dtype=np.float64)
        flat[:count * width] = np.asarray(rows, dtype=np.float64).reshape(-1)
        flat[cap * width:] = tail
        send = torch.from_numpy(flat)
        if dist.get_backend() == "nccl":
            send = send.cuda()
    recv = torch.empty(world * cap * (width + 2), dtype=torch.float64, device=send.device)
    dist.all_gather_into_tensor(recv, send)
    allr = recv.cpu().numpy().reshape(world, cap * (width + 2))
    out_rows, out_status, out_iters = [], [], []
    for r in range(world):
        n = shard_range(total, r, world)[1]
        out_rows.append(allr[r, :n * width].reshape(n, width))
        out_status.append(allr[r, cap * width:cap * width + n])
        out_iters.append(allr[r, cap * width + cap:cap * width + cap + n])
    return (np.concatenate(out_rows, axis=0), np.concatenate(out_status).astype(np.int32),
            np.concatenate(out_iters).astype(np.int32))

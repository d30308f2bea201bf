"""Multi-GPU plumbing for the sampling path: one process per GPU, image batches shard
embarrassingly, NCCL is used ONLY for the initial weight broadcast and the start/end barriers
(replaces the 8 redundant checkpoint loads of sample_c2i_ddp.py:58,72)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun-style init (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend)
    return rank, world, local


def broadcast_module(module: torch.nn.Module, src: int = 0) -> int:
    """Broadcast every parameter and buffer from `src` (one flat bucket per dtype so NVLink/NVSwitch sees
    a few large transfers instead of hundreds of small ones). Returns the number of bytes broadcast."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    by_dtype = {}
    for t in list(module.parameters()) + list(module.buffers()):
        by_dtype.setdefault(t.dtype, []).append(t.data)
    total = 0
    for dtype, tensors in by_dtype.items():
        flat = torch.cat([t.reshape(-1) for t in tensors])
        dist.broadcast(flat, src=src)
        off = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        total += flat.numel() * flat.element_size()
    return total


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `total` independent images for this rank (remainder to the low ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(global_seed: int, rank: int, world: int) -> int:
    return global_seed * world + rank          # sample_c2i_ddp.py:47


def image_index(i: int, rank: int, world: int, total_so_far: int) -> int:
    return i * world + rank + total_so_far     # sample_c2i_ddp.py:147

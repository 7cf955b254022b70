"""Multi-GPU plumbing: one process per GPU, batch sharding, one weight broadcast.

The inference path has no cross-sample dependency (InstanceNorm per sample, BatchNorm in eval
mode, per-sample volume / lookup / GRU), so the only collective is a single NCCL broadcast of
the weights at start-up.  Nothing is exchanged in steady state.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n_items for `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@torch.no_grad()
def broadcast_weights(module: torch.nn.Module, src: int = 0) -> int:
    """One flat broadcast of every parameter and buffer from `src`.  Returns the bytes sent."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    tensors = [t for t in list(module.parameters()) + list(module.buffers()) if t.is_floating_point()]
    others = [t for t in module.buffers() if not t.is_floating_point()]
    if not tensors:
        return 0
    flat = torch.cat([t.detach().reshape(-1).float() for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
        off += n
    for t in others:
        dist.broadcast(t, src=src)
    return flat.numel() * 4


def all_reduce_max(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shutdown() -> None:
    """Tear the process group down (silences NCCL's leak warning at interpreter exit)."""
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def all_gather_int(value: int, device) -> list:
    """Every rank's integer, in rank order (a list of one element without a process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(value)]
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]

"""Batch sharding across GPUs: one process per GPU, no data-path collective.

Every QP is independent, so rank ``r`` of ``W`` simply owns the
contiguous range ``[B*r/W, B*(r+1)/W)`` of the batch and its warm-start state;
``torch.distributed`` is only used for the barrier and the max-over-ranks time
of the benchmark and, optionally, to gather results on rank 0.
"""
from __future__ import annotations

import os
from typing import Tuple


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most 1) range of ``batch`` owned by ``rank``."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("need 0 <= rank < world")
    return batch * rank // world, batch * (rank + 1) // world


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment, (0, 0, 1) outside it."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def init_process_group(backend: str | None = None):
    """Join the torchrun job if there is one; returns (rank, local_rank, world)."""
    rank, local_rank, world = env_rank_world()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def max_over_ranks(value: float, device=None) -> float:
    """All-reduce MAX of a python float (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def gather_rows(local, batch: int):
    """Gather per-rank row blocks (numpy, shard_range order) onto every rank."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = np.zeros((mx,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.as_tensor(pad, device=dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:s] for o, s in zip(outs, sizes)], axis=0)

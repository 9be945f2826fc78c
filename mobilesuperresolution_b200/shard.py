"""Collective-free batch sharding across the GPUs of one box.

Images, patches and clips are independent and the weights are tiny (0.2-6 M params, replicated), so the
multi-GPU story of this path is: one process per GPU, each takes a contiguous slice of the batch, no data-path
collective (NVLink idle).  ``torch.distributed`` is only used for the rendezvous, a start barrier and the
max-over-ranks reduction of the measured time.  A clip's time axis is never split (the BasicVSR recurrence is
sequential, models/basicvsr_arch_origin.py:64-94).
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_slice(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of ``total`` items for ``rank``; sizes differ by at most one, earlier ranks larger."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def init_distributed(backend: str) -> Tuple[int, int, int]:
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, init_method="env://", rank=rank, world_size=world)
    return rank, local_rank, world


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: str = "cpu") -> float:
    """Slowest rank's value (multi-GPU timings are reported as the max over ranks)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: str = "cpu") -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())

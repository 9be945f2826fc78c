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


def _replicate(model):
    """copy.deepcopy of a module tree that holds old-style ``weight_norm`` convolutions: their hook leaves a computed, non-leaf
    ``weight`` attribute behind that deepcopy refuses; this package never reads it (filters are folded from weight_g / weight_v)."""
    import copy
    stripped = []
    for mod in model.modules():
        w = mod.__dict__.get("weight")
        if isinstance(w, torch.Tensor) and not w.is_leaf:
            stripped.append((mod, w))
            del mod.__dict__["weight"]
    try:
        return copy.deepcopy(model)
    finally:
        for mod, w in stripped:
            mod.__dict__["weight"] = w


class DeviceShards:
    """Batch sharding across the GPUs of one box from ONE process (the multi-process form is ``bench.py --gpus N`` under torchrun).

    One replica of ``model`` per device (parameters are 0.2-6 M: replicated), each call splits dim 0 of the host batch into contiguous
    shards (``shard_slice``), and every device runs H2D copy -> forward -> D2H copy on its own stream; nothing is exchanged between
    devices.  For video models dim 0 is the clip batch ``b``: a clip's time axis is never split.  The reference has no counterpart
    (its only multi-GPU mode is training DDP).

        runner = DeviceShards(model, precision="bf16")              # all visible GPUs
        y = runner(x_pinned)                                         # (N, 3, H, W) host tensor -> (N, 3, sH, sW) host tensor
        y = runner(clips_pinned, 720, 1280)                          # extra forward arguments are passed through
    """

    def __init__(self, model, devices=None, precision: str = "bf16", freeze: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceShards needs CUDA devices -- there is no CPU fallback")
        devs = list(range(torch.cuda.device_count())) if devices is None else list(devices)
        if not devs:
            raise ValueError("no devices")
        self.devices = [torch.device("cuda", int(d)) if not isinstance(d, torch.device) else d for d in devs]
        self.precision = precision
        self.replicas, self.streams = [], []
        for d in self.devices:
            r = _replicate(model).to(d).eval()
            if hasattr(r, "set_precision"):
                r.set_precision(precision)
            if freeze and hasattr(r, "freeze"):
                r.freeze(d)                      # fold + upload now, skip the per-call parameter walk
            self.replicas.append(r)
            self.streams.append(torch.cuda.Stream(device=d))

    def slices(self, total: int):
        return [shard_slice(total, r, len(self.devices)) for r in range(len(self.devices))]

    @torch.no_grad()
    def __call__(self, x_host: torch.Tensor, *fargs, out: "torch.Tensor | None" = None, in_dtype: "torch.dtype | None" = None):
        """``x_host``: CPU tensor (pinned memory makes the copies asynchronous).  Returns the CPU result (pinned when allocated here)."""
        if x_host.is_cuda:
            raise RuntimeError("DeviceShards takes the HOST batch and shards it; for a resident batch call the model directly")
        in_dtype = in_dtype or (torch.bfloat16 if self.precision == "bf16" and x_host.dtype == torch.float32 else x_host.dtype)
        parts = []
        for (lo, hi), d, s, m in zip(self.slices(x_host.shape[0]), self.devices, self.streams, self.replicas):
            if hi == lo:
                parts.append(None)
                continue
            with torch.cuda.device(d), torch.cuda.stream(s):
                xd = x_host[lo:hi].to(d, non_blocking=True).to(in_dtype)
                y = m(xd, *fargs)
                y = y[0] if isinstance(y, tuple) else y
                if out is None:
                    out = torch.empty((x_host.shape[0],) + tuple(y.shape[1:]), dtype=y.dtype, pin_memory=True)
                out[lo:hi].copy_(y, non_blocking=True)
                parts.append(y)                  # keep the device tensor alive until its copy has run
        for s in self.streams:
            s.synchronize()
        return out

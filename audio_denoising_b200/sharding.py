"""Data-parallel sharding of independent clips over the GPUs of one node.

Clips never interact (per-clip peak, per-clip GRU state, per-clip Griffin-Lim), so the hot path needs no
collective: rank r of W takes a contiguous block of clips and writes its own output slab.  torch.distributed is
used only off the data path (barrier + reducing timers / counters for reporting)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most 1) block of [0, n_items) owned by ``rank``."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _device_for_backend() -> torch.device:
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def max_over_ranks(value: float) -> float:
    """Step time of the job = slowest rank (reporting only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(value: int) -> int:
    """Sum of per-rank item counters (reporting only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(value)
    t = torch.tensor([value], dtype=torch.int64, device=_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())

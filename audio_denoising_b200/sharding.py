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


def bind_host_to_gpu(device_index: int) -> dict:
    """Pin the calling process to the CPUs next to GPU ``device_index`` (NVML's affinity mask for the device, intersected
    with what the container allows) so that the pinned staging buffers it allocates afterwards (first-touch) and the
    threads that feed them sit on the GPU's own NUMA node.  One process per GPU feeding 130 MB per step each way is
    host-memory bound; without the binding every rank's buffers land wherever the launcher happened to run.
    Returns what was done (for logging); never raises: a box without NVML or with a single node is left untouched."""
    import os

    info = {"bound": False}
    if os.environ.get("B2D_NO_NUMA_BIND"):
        return info
    try:
        import pynvml

        pynvml.nvmlInit()
        prop = torch.cuda.get_device_properties(device_index)
        bus_id = "%08x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        allowed = os.sched_getaffinity(0)
        words = (max(allowed | {os.cpu_count() or 1}) + 64) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        near = {w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus = sorted(near & allowed)
        info.update(gpu_bus=bus_id, near_cpus=len(near), allowed_cpus=len(allowed))
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=f"{cpus[0]}-{cpus[-1]} ({len(cpus)})")
    except Exception as e:  # noqa: BLE001 - best effort by design
        info["error"] = f"{type(e).__name__}: {e}"
    return info

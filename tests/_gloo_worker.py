"""2-rank gloo worker for tests/test_host_cpu.py: exercises the multi-process plumbing of the sharded driver
(rank -> clip range, per-rank result slab, max-over-ranks timing) without a GPU."""
import os
import time

import torch
import torch.distributed as dist

from audio_denoising_b200.sharding import gather_counts, max_over_ranks, shard_range

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_clips = 11
lo, hi = shard_range(n_clips, world, rank)
t0 = time.perf_counter()
local = torch.arange(lo, hi, dtype=torch.float32) * 2.0  # stand-in for the per-rank denoise of clips [lo, hi)
ms = (time.perf_counter() - t0) * 1e3 + rank  # rank 1 is "slower"
total = gather_counts(hi - lo)
worst = max_over_ranks(ms)
if rank == 0:
    assert total == n_clips and worst >= 1.0
    print(f"GLOO_OK clips={total} max_ms={worst:.3f}", flush=True)
dist.destroy_process_group()

"""Multi-GPU plumbing for the hot path: it shards by image pair and needs no data-path collective.

Every image pair is independent and the 25 MB of weights are replicated, so N GPUs = N replicas, each taking a
contiguous slice of the global batch.  The single collective is the gather of the per-sample metric
(the reference reduces metrics on the host after a .item() per step, runtime.py:438-459)."""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of the global batch owned by ``rank``; ragged batches give the first
    ``global_batch % world`` ranks one extra sample, empty slices are allowed."""
    if world < 1 or not (0 <= rank < world) or global_batch < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_metric(local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """All-gather per-sample metric values (1-D, one per local sample) into the global order.  Uses the initialised
    torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests); with no process group it is
    the identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(global_batch, world, r) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad = torch.zeros(width, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:b - a] for o, (a, b) in zip(out, sizes)])


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing rule: a multi-GPU step time is the MAX over ranks."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])

"""Batch sharding across the GPUs of one box: one process per GPU, weights replicated, no collective on the hot path.

Every sample's forward is independent (LayerNorm only, eval mode), so rank r of G simply takes a contiguous slice of
the batch. The only communication is the *optional* all-gather of the output embeddings (NCCL over NVLink on GPUs,
gloo in the CPU tests); it is outside the timed region of bench.py.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [start, end) of ``n`` samples for ``rank``; the first ``n % world`` ranks get one extra."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(x: Tensor, rank: int | None = None, world: int | None = None) -> Tensor:
    """This rank's slice of a batch-first tensor (a view, no copy)."""
    if rank is None or world is None:
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_embeddings(local: Tensor, total: int | None = None) -> Tensor:
    """All-gather per-rank outputs (rows may differ by one between ranks) back into batch order on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    if total is None:
        counts = torch.tensor([local.shape[0]], device=local.device)
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts)
        sizes = [int(c.item()) for c in all_counts]
    else:
        sizes = [hi - lo for lo, hi in (shard_bounds(total, r, world) for r in range(world))]
    width = max(sizes)
    padded = local
    if local.shape[0] < width:
        pad = torch.zeros(width - local.shape[0], *local.shape[1:], device=local.device, dtype=local.dtype)
        padded = torch.cat([local, pad])
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous())
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])

"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the batch sharding and the optional all-gather."""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytorch_models_b200.sharding import gather_embeddings, shard_batch, shard_bounds


def test_shard_bounds_cover_the_batch_exactly():
    for n in (0, 1, 7, 8, 128, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n * 4, dtype=torch.float32).reshape(n, 4)
        mine = shard_batch(full)
        lo, hi = shard_bounds(n, rank, world)
        assert torch.equal(mine, full[lo:hi])
        # stand-in for the per-rank forward: any per-sample function keeps batch order after the gather
        out = gather_embeddings(mine * 2 + 1, total=n)
        assert torch.equal(out, full * 2 + 1)
        out2 = gather_embeddings(mine * 2 + 1)  # sizes discovered with a small all-gather
        assert torch.equal(out2, full * 2 + 1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7])
def test_two_rank_shard_and_gather_gloo(n):
    mp.spawn(_worker, args=(2, _free_port(), n), nprocs=2, join=True)

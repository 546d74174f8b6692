"""CUDA-graph replay of a forward pass (launch-bound regime: small batches).

A ViT-B/16 forward is 65 kernel launches; below ~64 images the per-launch enqueue time of the Python host layer
(~1.9 ms; 0.25 ms through a launch plan, plans.py) exceeds the GPU time, and even where it does not the GPU itself
runs a captured forward faster than 65 stream launches (4.44 vs 4.75 ms at 128 images, round 2). All libb200enc entry
points are capturable (no host synchronisation, tensor maps passed by value, workspaces from
the caching allocator), so the whole forward can be captured once per input shape and replayed.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import ops, plans


class GraphedForward:
    """``g = GraphedForward(model, example)``; ``y = g(x)`` replays the captured forward on ``x`` (same shape/dtype).

    The packed-weight caches are built during the warm-up calls, so the graph reads the kernel-ready copies; after
    changing weights call :meth:`capture` again.
    """

    def __init__(self, module: nn.Module, example: Tensor, warmup: int = 2) -> None:
        if not example.is_cuda:
            raise RuntimeError("GraphedForward needs a CUDA example input")
        self.module = module
        self.static_in = example.detach().clone()
        self.capture(warmup)

    @torch.no_grad()
    def capture(self, warmup: int = 2) -> None:
        side = torch.cuda.Stream(self.static_in.device)
        side.wait_stream(torch.cuda.current_stream())
        prev = plans.enable(False)  # the warm-up runs on a side stream: no point in recording a launch plan for it
        try:
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self.module(self.static_in)
        finally:
            plans.enable(prev)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCHES
        with torch.cuda.graph(self.graph):
            self.static_out = self.module(self.static_in)
        self.launches = ops.LAUNCHES - n0  # libb200enc kernels inside the graph (counted on every replay)

    @torch.no_grad()
    def __call__(self, x: Tensor) -> Tensor:
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise ValueError(f"captured for {tuple(self.static_in.shape)} {self.static_in.dtype}, "
                             f"got {tuple(x.shape)} {x.dtype}")
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        ops.LAUNCHES += self.launches
        return self.static_out.clone()

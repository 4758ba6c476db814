"""CUDA-graph replay of one forward+backward of the JSD estimator.

The hot path is a short, fixed sequence of library kernels (and, with ``gather=True``, two NCCL
collectives); at B <= a few thousand it is bounded by launch / Python overhead rather than by the
GPU.  ``GraphedStep`` captures ``loss = loss_fn(f, g, t); grads = autograd.grad(loss, (f, g, t))``
once into a ``torch.cuda.CUDAGraph`` over static input buffers and replays it with a single launch
(SURVEY 8-f "next" #4).  Shapes, dtypes and the temperature tensor are fixed at capture time; the
caller copies new embeddings into ``.f`` / ``.g`` (or passes them to ``__call__``) before a replay.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch


class GraphedStep:
    def __init__(self, loss_fn: Callable, f: torch.Tensor, g: torch.Tensor, t: torch.Tensor, warmup: int = 3):
        if not f.is_cuda:
            raise RuntimeError("GraphedStep needs CUDA tensors")
        self.t = t
        self.f = f.detach().clone().requires_grad_(True)
        self.g = g.detach().clone().requires_grad_(True)

        def step():
            loss = loss_fn(self.f, self.g, self.t)
            loss = loss[0] if isinstance(loss, tuple) else loss
            return (loss,) + tuple(torch.autograd.grad(loss, (self.f, self.g, self.t)))

        side = torch.cuda.Stream(device=f.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                  # warm-up off the capture stream (library init, NCCL, allocator)
            for _ in range(max(warmup, 1)):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.df, self.dg, self.dt = step()

    def __call__(self, f: torch.Tensor = None, g: torch.Tensor = None) -> Tuple[torch.Tensor, ...]:
        """Replay; returns (loss, dF, dG, dt) -- static tensors that the next replay overwrites."""
        with torch.no_grad():
            if f is not None:
                self.f.copy_(f, non_blocking=True)
            if g is not None:
                self.g.copy_(g, non_blocking=True)
        self.graph.replay()
        return self.loss, self.df, self.dg, self.dt

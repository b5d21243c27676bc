"""CUDA-graph capture of a whole step (forward + losses + backward) that runs through the library.

At training batch sizes (16 pairs per GPU) a DeepFNet step is ~400 kernel launches of a few microseconds each: the GPU
waits for the host.  Every launch of the library goes to the current stream through the C ABI, allocates only through
PyTorch's caching allocator and never synchronises, so the step can be recorded once and replayed: ``GraphedStep`` does
what torch.cuda.make_graphed_callables does for plain modules, for an arbitrary callable with dict / tuple outputs.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    """Record ``fn(*inputs)`` (it may call .backward() on a loss; gradients must accumulate into pre-existing .grad
    tensors, e.g. fepe_b200.dist.FlatGradients) into one CUDA graph.  ``replay(*inputs)`` copies new inputs into the
    static buffers, replays and returns the static outputs (valid until the next replay).

    `fn` must be free of host synchronisation (.item(), .cpu(), torch.tensor(list, device=...)) and data-dependent shapes.
    """

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 2):
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # The split (hi / lo fp16) copies of the weights are cached per parameter version.  The capture must RECORD the
        # kernels that make them (a hit on a warm-up entry would bake pointers to memory the graph does not own, holding
        # weights that are stale after the first optimiser step), so the cache is emptied first; entries made during
        # the capture live in the graph's private pool, are rewritten by every replay, and must not be served to eager
        # callers afterwards, so it is emptied again.
        from . import mlp32
        mlp32._SPLIT_CACHE.clear()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)
        mlp32._SPLIT_CACHE.clear()

    def replay(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src)
        self.graph.replay()
        return self.static_out

"""Multi-GPU plumbing: one process per GPU, image pairs sharded contiguously, no collective on the
data path (pairs are independent -- SURVEY.md 8e).  The only exchange is the training-time gradient
all-reduce (the reference uses nn.DataParallel's reduce-to-GPU0, deepFEPE/train_good.py:309-314);
here it is ONE flattened all-reduce over NCCL/NVLink (gloo on CPU in the tests)."""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_range(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the batch owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a python float over all ranks (device timings are reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_mean_grads_(params: Iterable[torch.nn.Parameter], extra: torch.Tensor = None):
    """Average the gradients of `params` (and optionally the entries of `extra`, e.g. loss / metric
    sums) over all ranks with a single flattened all-reduce; writes the result back in place."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return extra
    grads = [p.grad for p in params if p.grad is not None]
    flat = [g.reshape(-1) for g in grads]
    if extra is not None:
        flat.append(extra.reshape(-1).to(flat[0].dtype) if flat else extra.reshape(-1))
    buf = torch.cat(flat)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    buf /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(buf[off:off + n].view_as(g))
        off += n
    if extra is not None:
        return buf[off:off + extra.numel()].view_as(extra).to(extra.dtype)
    return None

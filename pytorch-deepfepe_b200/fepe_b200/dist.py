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


class FlatGradients:
    """Gradients of `params` as views of ONE flat fp32 buffer (plus `extra_numel` trailing floats for loss / metric
    sums): autograd accumulates into the views in place, so the training-step exchange is a single all-reduce of
    `flat` -- no torch.cat before it, no per-parameter copy after it -- and every rank always reduces the same layout
    (a parameter that received no gradient on a rank contributes zeros instead of changing the message length).

    Use ``zero_()`` instead of ``optimizer.zero_grad(set_to_none=True)`` (which would drop the views; zero_grad with
    ``set_to_none=False`` is fine too).  Reference behaviour replaced: nn.DataParallel's reduce of the replica
    gradients onto GPU 0 (deepFEPE/train_good.py:309-314)."""

    ALIGN = 32          # floats: every view starts on a 128-byte boundary (the weight-gradient kernels store 16-byte vectors)

    def __init__(self, params: Iterable[torch.nn.Parameter], extra_numel: int = 0, fuse_accumulation: bool = False):
        """fuse_accumulation=True additionally marks the parameters as gradient SINKS: backward kernels of the library
        that accumulate anyway (the weight gradients of the tensor-core MLP: fepe_mlp32_wgrad / _first_bwd / _last_bwd /
        _affine_grads) then add straight into these views and hand autograd no gradient for them -- no temporary, no
        `grad += g` kernel per parameter and use (110 of them in a depth-5 DeepFNet step).  Only for training loops that
        read gradients from .grad after backward() (not torch.autograd.grad, no per-parameter hooks)."""
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradients: no trainable parameters")
        dev = self.params[0].device
        offs, n = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatGradients: parameters must be fp32 on one device")
            offs.append(n)
            n += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(n + extra_numel, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, offs):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            if fuse_accumulation:
                p._fepe_grad_sink = True
        self.extra = self.flat[n:]
        self.numel = n

    def zero_(self):
        self.flat.zero_()

    def check_views(self) -> bool:
        """True while every parameter's .grad still aliases the flat buffer (zero_grad(set_to_none=True) breaks it)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)

    def allreduce_mean_(self):
        """Average gradients (and `extra`) over all ranks in place: ONE collective on the flat buffer."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self.extra
        if not self.check_views():
            raise RuntimeError("FlatGradients: a parameter's .grad no longer aliases the flat buffer "
                               "(use FlatGradients.zero_() or zero_grad(set_to_none=False))")
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / dist.get_world_size())
        return self.extra


def allreduce_mean_grads_(params: Iterable[torch.nn.Parameter], extra: torch.Tensor = None):
    """Average the gradients of `params` (and optionally the entries of `extra`, e.g. loss / metric
    sums) over all ranks with a single flattened all-reduce; writes the result back in place.  Parameters whose
    .grad is None contribute zeros, so the message has the same layout on every rank (a rank on which a sub-network
    received no gradient can neither hang the collective nor shift other parameters' gradients).  Prefer
    FlatGradients, which needs no flatten / copy-back."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return extra
    params = [p for p in params if p.requires_grad]
    flat = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params]
    if extra is not None:
        flat.append(extra.reshape(-1).to(flat[0].dtype) if flat else extra.reshape(-1))
    buf = torch.cat(flat)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    buf /= dist.get_world_size()
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = buf[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(buf[off:off + n].view_as(p))
        off += n
    if extra is not None:
        return buf[off:off + extra.numel()].view_as(extra).to(extra.dtype)
    return None

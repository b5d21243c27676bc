"""Tensor-level entry points over the C ABI (device pointers + current CUDA stream).

PyTorch is used for device memory, streams and autograd plumbing only; all arithmetic of the
path happens in libfepe_b200.so.  Inputs must be CUDA fp32 tensors; anything else raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

IDENTITY_AFFINE = (1.0, 0.0, 1.0, 0.0)


def hw_affine(image_size) -> Tuple[float, float, float, float]:
    """(ax,bx,ay,by) of the reference's NormalizeAndExpand_HW (deepFEPE/models/DeepFNet.py:111)."""
    H, W = float(image_size[0]), float(image_size[1])
    return (2.0 / W, -1.0, 2.0 / H, -1.0)


def _check_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"fepe_b200: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"fepe_b200: {name} must be float32, got {t.dtype}")
    return t.contiguous()


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def fit_forward(matches: torch.Tensor, weights: torch.Tensor, affine=IDENTITY_AFFINE,
                clamp_at: float = 0.5, want_epi: bool = True, want_saved: bool = False):
    """matches [B,N,4], weights [B,N] (or [B,1,N]) -> F [B,3,3], residual [B,N], epi [B,N]|None,
    saved [B,64] float64|None.  One launch of the fused kernel (include/fepe_b200.h: fepe_fit_fwd)."""
    matches = _check_cuda_f32(matches, "matches")
    weights = _check_cuda_f32(weights, "weights")
    B, N, four = matches.shape
    if four != 4:
        raise RuntimeError("fepe_b200: matches must be [B,N,4] (x1,y1,x2,y2)")
    weights = weights.reshape(B, N)
    dev = matches.device
    with torch.cuda.device(dev):
        F = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
        res = torch.empty(B, N, dtype=torch.float32, device=dev)
        epi = torch.empty(B, N, dtype=torch.float32, device=dev) if want_epi else None
        saved = torch.empty(B, _lib.SAVED_DOUBLES, dtype=torch.float64, device=dev) if want_saved else None
        st = _lib.lib().fepe_fit_fwd(matches.data_ptr(), weights.data_ptr(), B, N,
                                     affine[0], affine[1], affine[2], affine[3], float(clamp_at),
                                     F.data_ptr(), res.data_ptr(),
                                     epi.data_ptr() if epi is not None else None,
                                     saved.data_ptr() if saved is not None else None,
                                     _stream_ptr())
    _lib.check(st, "fepe_fit_fwd")
    return F, res, epi, saved

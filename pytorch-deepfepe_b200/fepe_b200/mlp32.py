"""fp32-parity tensor-core path of ErrorEstimator: drives the fepe_mlp32_* entry points of the C ABI
(include/fepe_b200.h).  Activations and parameters are fp32 in memory; the GEMMs run on tcgen05 with every operand
split into an fp16 (hi, lo) pair and three MMAs per product (csrc/fepe_mlp32.cu), fp32 accumulation, fp64 statistics.
This is the DEFAULT path of the weight networks (the reference computes them in fp32,
deepFEPE/models/ErrorEstimators.py:46-64); the bf16 path of mlp_tc.py is the opt-in fast mode.

`MLP32.__call__(matches, affine, extras)` evaluates the network straight from the model's inputs -- the pixel matches
and the channel groups the reference concatenates (DeepFNet.get_input :359-404, torch.cat :487) -- or from a plain
feature tensor.  Output: logits [B, out, N] and, for a one-logit network, the softmax over N.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import _lib

_CH = (64, 128, 1024, 512, 256)
SLOPE = 0.01                         # nn.LeakyReLU() default, ErrorEstimators.py:51
MAX_CIN = 16                         # fepe_mlp32_first


def _layers(fw: nn.Sequential):
    convs = [m for m in fw if isinstance(m, nn.Conv1d)]
    norms = [m for m in fw if isinstance(m, nn.InstanceNorm1d)]
    if [c.out_channels for c in convs[:5]] != list(_CH) or len(convs) != 6 or len(norms) != 5:
        raise RuntimeError("fepe_b200.mlp32 supports the reference ErrorEstimator layout (64/128/1024/512/256 -> out)")
    if convs[5].out_channels not in (1, 4):
        raise RuntimeError("fepe_b200.mlp32: the last layer must have 1 (weights) or 4 (offsets) outputs")
    if convs[0].in_channels > MAX_CIN:
        raise RuntimeError(f"fepe_b200.mlp32: at most {MAX_CIN} input channels")
    return convs, norms


def split_weight(lib, w2d: torch.Tensor, st):
    """W [Co,K] fp32 -> (Whi, Wlo fp16, wscale [4] fp32) through fepe_mlp32_prepare_weights."""
    Co, K = w2d.shape
    whi = torch.empty(Co, K, dtype=torch.float16, device=w2d.device)
    wlo = torch.empty_like(whi)
    wsc = torch.empty(4, dtype=torch.float32, device=w2d.device)
    _lib.check(lib.fepe_mlp32_prepare_weights(w2d.data_ptr(), whi.data_ptr(), wlo.data_ptr(), wsc.data_ptr(), Co, K, st),
               "fepe_mlp32_prepare_weights")
    return whi, wlo, wsc


def first_layer_args(matches, affine, extras, cin: int):
    """ctypes arguments of fepe_mlp32_first for the channel groups; returns (args, tensors to keep alive)."""
    keep = []
    total = 0
    if matches is not None:
        m = matches.contiguous()
        keep.append(m)
        mp, total = m.data_ptr(), 4
        ax, bx, ay, by = (float(v) for v in affine)
    else:
        mp, (ax, bx, ay, by) = None, (1.0, 0.0, 1.0, 0.0)
    ex = []
    for e in extras:
        e = e.contiguous()
        keep.append(e)
        c = 1 if e.dim() == 2 else e.shape[2]
        ex += [e.data_ptr(), c]
        total += c
    if len(extras) > 4:
        raise RuntimeError("fepe_b200.mlp32: at most four extra channel groups")
    ex += [None, 0] * (4 - len(extras))
    if total != cin:
        raise RuntimeError(f"fepe_b200.mlp32: the inputs carry {total} channels, the network expects {cin}")
    return [mp, ax, bx, ay, by] + ex, keep


class MLP32:
    """Inference (torch.no_grad()) evaluation of one ErrorEstimator."""

    def __init__(self, fw: nn.Sequential):
        self.convs, self.norms = _layers(fw)
        self.fw = fw
        self.cin = self.convs[0].in_channels
        self.cout = self.convs[5].out_channels
        self._versions = None

    def _refresh(self, lib, st):
        vers = tuple(p._version for p in self.fw.parameters()) + (next(self.fw.parameters()).device,)
        if vers == self._versions:
            return
        c, n = self.convs, self.norms
        self.w0 = c[0].weight.detach().reshape(64, self.cin).float().contiguous()
        self.w = [None] + [split_weight(lib, m.weight.detach().reshape(m.out_channels, m.in_channels).float().contiguous(), st)
                           for m in c[1:5]]
        self.w_last = c[5].weight.detach().reshape(self.cout, 256).float().contiguous()
        self.b_last = c[5].bias.detach().float().contiguous()
        self.gamma = [m.weight.detach().float().contiguous() for m in n]
        self.beta = [m.bias.detach().float().contiguous() for m in n]
        self.eps = [float(m.eps) for m in n]
        self._versions = vers

    def __call__(self, matches: Optional[torch.Tensor], affine, extras: Sequence[torch.Tensor], B: int, N: int):
        """matches [B,N,4] fp32 cuda or None; extras: channel groups [B,N,c] / [B,N] in the reference's order.
        Returns (logits [B,out,N], softmax weights [B,1,N] | None)."""
        ref = matches if matches is not None else extras[0]
        if not ref.is_cuda or ref.dtype != torch.float32:
            raise RuntimeError("fepe_b200.mlp32 needs CUDA fp32 inputs (there is no CPU path)")
        dev = ref.device
        lib = _lib.lib()
        Npad = (N + 127) // 128 * 128
        M = B * Npad
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            self._refresh(lib, st)
            bufa = torch.empty(M * 1024, dtype=torch.float32, device=dev)
            bufb = torch.empty(M * 512, dtype=torch.float32, device=dev)
            stats = torch.zeros(B * 1024 * 2, dtype=torch.float64, device=dev)
            ss = torch.empty(B * 1024 * 2, dtype=torch.float32, device=dev)
            args, keep = first_layer_args(matches, affine, extras, self.cin)
            # conv biases in front of an InstanceNorm cancel in the mean subtraction: not applied (bias = NULL)
            _lib.check(lib.fepe_mlp32_first(*args, self.w0.data_ptr(), None, bufa.data_ptr(), stats.data_ptr(),
                                            B, N, Npad, 64, st), "fepe_mlp32_first")
            src, dst, k = bufa, bufb, 64
            for i in range(1, 5):
                co = _CH[i]
                _lib.check(lib.fepe_mlp32_scale_shift(stats.data_ptr(), self.gamma[i - 1].data_ptr(),
                                                      self.beta[i - 1].data_ptr(), ss.data_ptr(), B, k, N, self.eps[i - 1],
                                                      1, st), "fepe_mlp32_scale_shift")
                whi, wlo, wsc = self.w[i]
                _lib.check(lib.fepe_mlp32_gemm(src.data_ptr(), ss.data_ptr(), SLOPE, whi.data_ptr(), wlo.data_ptr(),
                                               wsc.data_ptr(), None, dst.data_ptr(), stats.data_ptr(), B, Npad, N, k, co,
                                               st), "fepe_mlp32_gemm")
                src, dst, k = dst, src, co
            _lib.check(lib.fepe_mlp32_scale_shift(stats.data_ptr(), self.gamma[4].data_ptr(), self.beta[4].data_ptr(),
                                                  ss.data_ptr(), B, 256, N, self.eps[4], 0, st), "fepe_mlp32_scale_shift")
            logits = torch.empty(B, self.cout, N, dtype=torch.float32, device=dev)
            weights = torch.empty(B, 1, N, dtype=torch.float32, device=dev) if self.cout == 1 else None
            _lib.check(lib.fepe_mlp32_last(src.data_ptr(), ss.data_ptr(), SLOPE, self.w_last.data_ptr(),
                                           self.b_last.data_ptr(), logits.data_ptr(),
                                           weights.data_ptr() if weights is not None else None, B, N, Npad, 256,
                                           self.cout, st), "fepe_mlp32_last")
            del keep
        return logits, weights

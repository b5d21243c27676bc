"""Tensor-core (tcgen05) inference path of ErrorEstimator: drives the fepe_mlp_* entry points of the
C ABI (include/fepe_b200.h).  bf16 activations and weights, fp32 accumulation and statistics.

The training path keeps PyTorch's fp32 kernels (autograd); this path is used under torch.no_grad()
when a module was switched on with ``ErrorEstimator.tensor_cores = True`` /
``DeepFNet.enable_tensor_core_mlp()``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

_CH = (64, 128, 1024, 512, 256)


class TensorCoreMLP:
    def __init__(self, fw: nn.Sequential):
        convs = [m for m in fw if isinstance(m, nn.Conv1d)]
        norms = [m for m in fw if isinstance(m, nn.InstanceNorm1d)]
        if [c.out_channels for c in convs] != list(_CH) + [1] or len(norms) != 5:
            raise RuntimeError("TensorCoreMLP supports the reference ErrorEstimator layout (64/128/1024/512/256 -> 1)")
        self.fw, self.convs, self.norms = fw, convs, norms
        self.cin = convs[0].in_channels
        if self.cin > 8:
            raise RuntimeError("TensorCoreMLP: the first layer supports at most 8 input channels")
        self._versions = None
        self._buf_key = None

    # -- parameters in the layouts the kernels want (refreshed when the module's parameters change) --
    def _refresh(self):
        vers = tuple(p._version for p in self.fw.parameters()) + (next(self.fw.parameters()).device,)
        if vers == self._versions:
            return
        c, n = self.convs, self.norms
        self.w0 = c[0].weight.detach().reshape(64, self.cin).float().contiguous()
        self.b = [m.bias.detach().float().contiguous() for m in c]
        self.w = [None] + [m.weight.detach().reshape(m.out_channels, m.in_channels).to(torch.bfloat16).contiguous()
                           for m in c[1:5]]
        self.w_last = c[5].weight.detach().reshape(256).float().contiguous()
        self.b_last = float(c[5].bias.detach().item())
        self.gamma = [m.weight.detach().float().contiguous() for m in n]
        self.beta = [m.bias.detach().float().contiguous() for m in n]
        self.eps = [float(m.eps) for m in n]
        self._versions = vers

    def _buffers(self, B, Npad, dev):
        key = (B, Npad, dev)
        if key != self._buf_key:
            self.act = [torch.empty(B * Npad, 1024, dtype=torch.bfloat16, device=dev) for _ in range(2)]
            self.stats = torch.empty(B, 1024, 2, dtype=torch.float32, device=dev)
            self._buf_key = key
        return self.act, self.stats

    def __call__(self, x: torch.Tensor):
        """x [B,Cin,N] fp32 cuda -> (logits [B,1,N], softmax weights [B,1,N]) fp32."""
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("TensorCoreMLP needs a CUDA fp32 input")
        self._refresh()
        B, Cin, N = x.shape
        Npad = (N + 127) // 128 * 128
        dev = x.device
        lib = _lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        x0 = x.permute(0, 2, 1).contiguous()
        (ya, xa), stats = self._buffers(B, Npad, dev)
        slope = 0.01
        with torch.cuda.device(dev):
            stats.zero_()
            _lib.check(lib.fepe_mlp_first(x0.data_ptr(), self.w0.data_ptr(), self.b[0].data_ptr(), ya.data_ptr(),
                                          stats.data_ptr(), B, N, Npad, Cin, 64, st), "fepe_mlp_first")
            _lib.check(lib.fepe_mlp_norm(ya.data_ptr(), stats.data_ptr(), self.gamma[0].data_ptr(),
                                         self.beta[0].data_ptr(), xa.data_ptr(), B, Npad, N, 64, self.eps[0], slope, st),
                       "fepe_mlp_norm")
            k = 64
            for i in range(1, 5):
                co = _CH[i]
                stats.zero_()
                _lib.check(lib.fepe_mlp_gemm(xa.data_ptr(), self.w[i].data_ptr(), self.b[i].data_ptr(), ya.data_ptr(),
                                             stats.data_ptr(), B, Npad, N, k, co, st), "fepe_mlp_gemm")
                _lib.check(lib.fepe_mlp_norm(ya.data_ptr(), stats.data_ptr(), self.gamma[i].data_ptr(),
                                             self.beta[i].data_ptr(), xa.data_ptr(), B, Npad, N, co, self.eps[i], slope,
                                             st), "fepe_mlp_norm")
                k = co
            logits = torch.empty(B, 1, N, dtype=torch.float32, device=dev)
            weights = torch.empty(B, 1, N, dtype=torch.float32, device=dev)
            _lib.check(lib.fepe_mlp_last(xa.data_ptr(), self.w_last.data_ptr(), self.b_last, logits.data_ptr(),
                                         weights.data_ptr(), B, N, Npad, 256, st), "fepe_mlp_last")
        return logits, weights

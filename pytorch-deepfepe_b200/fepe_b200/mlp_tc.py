"""Tensor-core (tcgen05) inference path of ErrorEstimator: drives the fepe_mlp_* entry points of the
C ABI (include/fepe_b200.h).  bf16 activations and weights, fp32 accumulation and statistics.

Inference (torch.no_grad()): ``TensorCoreMLP.__call__`` -- ping-pong activation buffers, softmax fused.
Training: ``TensorCoreMLPFunction`` (torch.autograd.Function) -- the same forward kernels keeping every
layer's pre-norm output Y and block output X', and a backward made of fepe_mlp_last_bwd, fepe_mlp_normbwd
(InstanceNorm + LeakyReLU adjoint), fepe_mlp_wgrad (MN-major tcgen05 GEMM, dW = dY^T X), the data-gradient
GEMM (fepe_mlp_gemm with W^T) and fepe_mlp_first_bwd.  Both are opt-in (``DeepFNet.set_mlp_path("bf16")``): the default
is the fp32-parity tensor-core path of fepe_b200/mlp32.py.
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
        # fuse_norm: InstanceNorm + LeakyReLU of layers 1-4 applied to the next GEMM's operand tiles in shared memory
        # (fepe_mlp_gemm_norm) instead of a separate pass over memory; same bf16 results as the unfused kernels.
        self.fuse_norm = True
        # A Conv1d bias in front of an InstanceNorm cancels in the mean subtraction; the inference GEMMs skip it (the
        # rounding of Y to bf16 is then relative to the centred scale of the channel, never to its offset).
        self.drop_bias = True
        self.unfused_k = (128,)         # operand widths whose layer keeps the separate norm kernel (see _fused_tail)

    # -- parameters in the layouts the kernels want (refreshed when the module's parameters change) --
    def _refresh(self):
        tensors = [t for m in self.convs + self.norms for t in (m.weight, m.bias)]       # (DataParallel replicas: see mlp32)
        vers = tuple((t._version, t.data_ptr()) for t in tensors) + (tensors[0].device,)
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
        # allocated per call through the caching allocator (stream ordered): a module evaluated from several streams or
        # threads (nn.DataParallel replicas share nothing, but user code may) must not share scratch buffers
        act = [torch.empty(B * Npad, 1024, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        stats = torch.empty(B, 1024, 2, dtype=torch.float32, device=dev)
        ss = torch.empty(B, 512, 4, dtype=torch.float32, device=dev)
        xn = torch.empty(B * Npad, 128, dtype=torch.bfloat16, device=dev)   # normalised operand of an unfused layer
        return act, stats, ss, xn

    def __call__(self, x: torch.Tensor):
        """x [B,Cin,N] fp32 cuda -> (logits [B,1,N], softmax weights [B,1,N]) fp32."""
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("TensorCoreMLP needs a CUDA fp32 input")
        self._refresh()
        B, Cin, N = x.shape
        Npad = (N + 127) // 128 * 128
        dev = x.device
        lib = _lib.lib()
        st = torch.cuda.current_stream(dev).cuda_stream
        x0 = x.permute(0, 2, 1).contiguous()
        (ya, xa), stats, ss, xn = self._buffers(B, Npad, dev)
        slope = 0.01
        with torch.cuda.device(dev):
            stats.zero_()
            _lib.check(lib.fepe_mlp_first(x0.data_ptr(), self.w0.data_ptr(), self.b[0].data_ptr(), ya.data_ptr(),
                                          stats.data_ptr(), B, N, Npad, Cin, 64, st), "fepe_mlp_first")
            k = 64
            bias = [0 if self.drop_bias else b.data_ptr() for b in self.b]
            if self.fuse_norm:
                return self._fused_tail(lib, st, ya, xa, stats, ss, xn, bias, B, N, Npad, dev, slope)
            _lib.check(lib.fepe_mlp_norm(ya.data_ptr(), stats.data_ptr(), self.gamma[0].data_ptr(),
                                         self.beta[0].data_ptr(), xa.data_ptr(), B, Npad, N, 64, self.eps[0], slope, st),
                       "fepe_mlp_norm")
            for i in range(1, 5):
                co = _CH[i]
                stats.zero_()
                _lib.check(lib.fepe_mlp_gemm(xa.data_ptr(), self.w[i].data_ptr(), bias[i], ya.data_ptr(),
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


    def _fused_tail(self, lib, st, src, dst, stats, ss, xn, bias, B, N, Npad, dev, slope):
        """Layers 2-5 with the previous layer's norm fused into the GEMM operand path; `src` holds layer 1's pre-norm
        output and `stats` its statistics."""
        k = 64
        for i in range(1, 5):
            co = _CH[i]
            if k in self.unfused_k:
                # thin operand, wide output (128 -> 1024): the GEMM is bound by its epilogue, the operand would be
                # transformed once per n-tile, and the norm pass over 128 channels is cheap -- measured faster unfused
                _lib.check(lib.fepe_mlp_norm(src.data_ptr(), stats.data_ptr(), self.gamma[i - 1].data_ptr(),
                                             self.beta[i - 1].data_ptr(), xn.data_ptr(), B, Npad, N, k,
                                             self.eps[i - 1], slope, st), "fepe_mlp_norm")
                stats.zero_()
                _lib.check(lib.fepe_mlp_gemm(xn.data_ptr(), self.w[i].data_ptr(), bias[i], dst.data_ptr(),
                                             stats.data_ptr(), B, Npad, N, k, co, st), "fepe_mlp_gemm")
            else:
                _lib.check(lib.fepe_mlp_scale_shift(stats.data_ptr(), self.gamma[i - 1].data_ptr(),
                                                    self.beta[i - 1].data_ptr(), ss.data_ptr(), B, k, N, self.eps[i - 1],
                                                    1, st), "fepe_mlp_scale_shift")
                _lib.check(lib.fepe_mlp_gemm_norm(src.data_ptr(), ss.data_ptr(), slope, self.w[i].data_ptr(), bias[i],
                                                  dst.data_ptr(), stats.data_ptr(), B, Npad, N, k, co, st),
                           "fepe_mlp_gemm_norm")
            src, dst = dst, src
            k = co
        _lib.check(lib.fepe_mlp_scale_shift(stats.data_ptr(), self.gamma[4].data_ptr(), self.beta[4].data_ptr(),
                                            ss.data_ptr(), B, 256, N, self.eps[4], 0, st), "fepe_mlp_scale_shift")
        logits = torch.empty(B, 1, N, dtype=torch.float32, device=dev)
        weights = torch.empty(B, 1, N, dtype=torch.float32, device=dev)
        _lib.check(lib.fepe_mlp_last_norm(src.data_ptr(), ss.data_ptr(), slope, self.w_last.data_ptr(), self.b_last,
                                          logits.data_ptr(), weights.data_ptr(), B, N, Npad, 256, st), "fepe_mlp_last_norm")
        return logits, weights


class TensorCoreMLPFunction(torch.autograd.Function):
    """logits = ErrorEstimator(x) with bf16 tensor-core GEMMs in forward AND backward.

    forward(x [B,Cin,N] fp32, w1,b1,g1,be1, ..., w5,b5,g5,be5, w6,b6) -> logits [B,1,N] fp32.
    Parameter order = the module's own (conv weight [Co,Ci,1], conv bias, norm weight, norm bias) x 5, then the
    last conv's weight and bias."""

    @staticmethod
    def forward(ctx, x, *params):
        lib = _lib.lib()
        B, Cin, N = x.shape
        Npad = (N + 127) // 128 * 128
        dev = x.device
        st = torch.cuda.current_stream(dev).cuda_stream
        M = B * Npad
        convw = [params[4 * i] for i in range(5)] + [params[20]]
        convb = [params[4 * i + 1] for i in range(5)] + [params[21]]
        gam = [params[4 * i + 2].detach().float().contiguous() for i in range(5)]
        bet = [params[4 * i + 3].detach().float().contiguous() for i in range(5)]
        x0 = x.detach().permute(0, 2, 1).contiguous().float()
        w0 = convw[0].detach().reshape(64, Cin).float().contiguous()
        wb = [None] + [convw[i].detach().reshape(_CH[i], _CH[i - 1]).to(torch.bfloat16).contiguous() for i in range(1, 5)]
        bias = [b.detach().float().contiguous() for b in convb]
        w_last = convw[5].detach().reshape(256).float().contiguous()
        Ys = [torch.empty(M, c, dtype=torch.bfloat16, device=dev) for c in _CH]
        Xs = [torch.empty(M, c, dtype=torch.bfloat16, device=dev) for c in _CH]
        stats = [torch.zeros(B, c, 2, dtype=torch.float32, device=dev) for c in _CH]
        eps, slope = 1e-5, 0.01
        with torch.cuda.device(dev):
            _lib.check(lib.fepe_mlp_first(x0.data_ptr(), w0.data_ptr(), bias[0].data_ptr(), Ys[0].data_ptr(),
                                          stats[0].data_ptr(), B, N, Npad, Cin, 64, st), "fepe_mlp_first")
            for i in range(5):
                if i > 0:
                    _lib.check(lib.fepe_mlp_gemm(Xs[i - 1].data_ptr(), wb[i].data_ptr(), bias[i].data_ptr(),
                                                 Ys[i].data_ptr(), stats[i].data_ptr(), B, Npad, N, _CH[i - 1], _CH[i],
                                                 st), "fepe_mlp_gemm")
                _lib.check(lib.fepe_mlp_norm(Ys[i].data_ptr(), stats[i].data_ptr(), gam[i].data_ptr(), bet[i].data_ptr(),
                                             Xs[i].data_ptr(), B, Npad, N, _CH[i], eps, slope, st), "fepe_mlp_norm")
            logits = torch.empty(B, 1, N, dtype=torch.float32, device=dev)
            scratch = torch.empty(B, 1, N, dtype=torch.float32, device=dev)
            # the bias of the last layer is added on the device (the ABI takes it as a scalar: reading it here would
            # synchronise with the host every step and bake a stale value into a captured graph)
            _lib.check(lib.fepe_mlp_last(Xs[4].data_ptr(), w_last.data_ptr(), 0.0, logits.data_ptr(),
                                         scratch.data_ptr(), B, N, Npad, 256, st), "fepe_mlp_last")
            logits.add_(bias[5].reshape(1, 1, 1))
        ctx.dims = (B, Cin, N, Npad)
        ctx.saved = (x0, w0, wb, w_last, gam, Ys, Xs, stats)
        ctx.param_shapes = [p.shape for p in params]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.lib()
        B, Cin, N, Npad = ctx.dims
        x0, w0, wb, w_last, gam, Ys, Xs, stats = ctx.saved
        dev = x0.device
        st = torch.cuda.current_stream(dev).cuda_stream
        M = B * Npad
        eps, slope = 1e-5, 0.01
        dl = dlogits.detach().reshape(B, N).float().contiguous()
        grads = [None] * 22
        zeros_b = lambda c: torch.zeros(c, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            dX = torch.empty(M, 256, dtype=torch.bfloat16, device=dev)
            dw6, db6 = zeros_b(256), zeros_b(1)
            _lib.check(lib.fepe_mlp_last_bwd(dl.data_ptr(), Xs[4].data_ptr(), w_last.data_ptr(), dX.data_ptr(),
                                             dw6.data_ptr(), db6.data_ptr(), B, N, Npad, 256, st), "fepe_mlp_last_bwd")
            grads[20], grads[21] = dw6.reshape(1, 256, 1), db6
            for i in range(4, -1, -1):
                c = _CH[i]
                A = torch.zeros(B, c, 2, dtype=torch.float32, device=dev)
                dY = torch.empty(M, c, dtype=torch.bfloat16, device=dev)
                _lib.check(lib.fepe_mlp_normbwd(dX.data_ptr(), Xs[i].data_ptr(), Ys[i].data_ptr(), stats[i].data_ptr(),
                                                gam[i].data_ptr(), A.data_ptr(), dY.data_ptr(), B, Npad, N, c, eps, slope,
                                                st), "fepe_mlp_normbwd")
                grads[4 * i + 2] = A[:, :, 1].sum(0)          # dgamma
                grads[4 * i + 3] = A[:, :, 0].sum(0)          # dbeta
                grads[4 * i + 1] = zeros_b(c)                  # conv bias before InstanceNorm: exactly zero gradient
                if i > 0:
                    ci = _CH[i - 1]
                    dW = torch.zeros(c, ci, dtype=torch.float32, device=dev)
                    _lib.check(lib.fepe_mlp_wgrad(dY.data_ptr(), Xs[i - 1].data_ptr(), dW.data_ptr(), M, c, ci, st),
                               "fepe_mlp_wgrad")
                    grads[4 * i] = dW.reshape(c, ci, 1)
                    wt = wb[i].t().contiguous()                # [ci, c]: the data-gradient GEMM's "weight"
                    dXn = torch.empty(M, ci, dtype=torch.bfloat16, device=dev)
                    _lib.check(lib.fepe_mlp_gemm(dY.data_ptr(), wt.data_ptr(), None, dXn.data_ptr(),
                                                 None, B, Npad, Npad, c, ci, st), "fepe_mlp_gemm(dgrad)")
                    dX = dXn
                else:
                    dW = torch.zeros(64, Cin, dtype=torch.float32, device=dev)
                    dx0 = torch.zeros(B, N, Cin, dtype=torch.float32, device=dev)
                    _lib.check(lib.fepe_mlp_first_bwd(dY.data_ptr(), x0.data_ptr(), w0.data_ptr(), dx0.data_ptr(),
                                                      dW.data_ptr(), B, N, Npad, Cin, 64, st), "fepe_mlp_first_bwd")
                    grads[0] = dW.reshape(64, Cin, 1)
                    gx = dx0.permute(0, 2, 1).contiguous()
        return (gx,) + tuple(grads)


def module_params(fw: nn.Sequential):
    """The 22 parameters of an ErrorEstimator in the order TensorCoreMLPFunction expects."""
    convs = [m for m in fw if isinstance(m, nn.Conv1d)]
    norms = [m for m in fw if isinstance(m, nn.InstanceNorm1d)]
    out = []
    for i in range(5):
        out += [convs[i].weight, convs[i].bias, norms[i].weight, norms[i].bias]
    return out + [convs[5].weight, convs[5].bias]

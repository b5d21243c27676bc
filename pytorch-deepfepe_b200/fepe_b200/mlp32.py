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

import ctypes

import weakref
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


_SPLIT_CACHE = {}


def split_param_cached(lib, param: torch.Tensor, co: int, k: int, transposed: bool, st):
    """(hi, lo, scale) of a Conv1d weight [co, k, 1] (or of its transpose [k, co]: the data-gradient operand), cached per
    parameter VERSION: the update network is evaluated depth-1 times per step with the same weights, forward and
    backward, so each split is computed once per optimizer step instead of 2 x (depth - 1) times."""
    key = (id(param), bool(transposed))
    ver = (param._version, param.data_ptr(), param.device)
    hit = _SPLIT_CACHE.get(key)
    # id() values are reused once an object dies: the entry must belong to THIS parameter object (weak reference)
    if hit is not None and hit[0] == ver and hit[2]() is param:
        return hit[1]
    w2d = param.detach().reshape(co, k).float()
    w2d = w2d.t().contiguous() if transposed else w2d.contiguous()
    out = split_weight(lib, w2d, st)
    if len(_SPLIT_CACHE) >= 32:                        # drop the entries of dead parameters (e.g. DataParallel replicas)
        for k in [k for k, v in _SPLIT_CACHE.items() if v[2]() is None]:
            del _SPLIT_CACHE[k]
        if len(_SPLIT_CACHE) > 256:
            _SPLIT_CACHE.clear()
    _SPLIT_CACHE[key] = (ver, out, weakref.ref(param))
    return out


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
        # (the layers' own tensors, not fw.parameters(): nn.DataParallel replicas carry no registered parameters)
        tensors = [t for m in self.convs + self.norms for t in (m.weight, m.bias)]
        vers = tuple((t._version, t.data_ptr()) for t in tensors) + (tensors[0].device,)
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
            _lib.check(lib.fepe_mlp32_first(*args, self.w0.data_ptr(), None, bufa.data_ptr(), stats.data_ptr(), None,
                                            B, N, Npad, 64, st), "fepe_mlp32_first")
            src, dst, k = bufa, bufb, 64
            for i in range(1, 5):
                co = _CH[i]
                _lib.check(lib.fepe_mlp32_scale_shift(stats.data_ptr(), self.gamma[i - 1].data_ptr(),
                                                      self.beta[i - 1].data_ptr(), ss.data_ptr(), None, B, k, N,
                                                      self.eps[i - 1], 1, st), "fepe_mlp32_scale_shift")
                whi, wlo, wsc = self.w[i]
                _lib.check(lib.fepe_mlp32_gemm(src.data_ptr(), ss.data_ptr(), SLOPE, None, whi.data_ptr(), wlo.data_ptr(),
                                               wsc.data_ptr(), None, dst.data_ptr(), stats.data_ptr(), B, Npad, N, k, co,
                                               st), "fepe_mlp32_gemm")
                src, dst, k = dst, src, co
            _lib.check(lib.fepe_mlp32_scale_shift(stats.data_ptr(), self.gamma[4].data_ptr(), self.beta[4].data_ptr(),
                                                  ss.data_ptr(), None, B, 256, N, self.eps[4], 0, st), "fepe_mlp32_scale_shift")
            logits = torch.empty(B, self.cout, N, dtype=torch.float32, device=dev)
            weights = torch.empty(B, 1, N, dtype=torch.float32, device=dev) if self.cout == 1 else None
            _lib.check(lib.fepe_mlp32_last(src.data_ptr(), ss.data_ptr(), SLOPE, self.w_last.data_ptr(),
                                           self.b_last.data_ptr(), logits.data_ptr(),
                                           weights.data_ptr() if weights is not None else None, B, N, Npad, 256,
                                           self.cout, st), "fepe_mlp32_last")
            del keep
        return logits, weights


class MLP32Function(torch.autograd.Function):
    """logits = ErrorEstimator(inputs) under autograd: the forward kernels of MLP32 keeping every block's pre-norm output
    y (fp32), its (a, d) and (mean, rstd); the backward is fepe_mlp32_last_bwd, fepe_mlp32_normbwd, fepe_mlp32_wgrad and
    the data-gradient GEMM on tensor cores (split-fp16 operands), fepe_mlp32_first_bwd -- no library kernel on the path.

    forward(meta, matches | None, *extras, *params) -> logits [B, out, N];
    meta = (affine, n_extras, B, N); params = module_params(fw) (22 tensors)."""

    @staticmethod
    def forward(ctx, meta, matches, *tensors):
        affine, n_extras, B, N = meta
        extras, params = tensors[:n_extras], tensors[n_extras:]
        lib = _lib.lib()
        ref = matches if matches is not None else extras[0]
        dev = ref.device
        Npad = (N + 127) // 128 * 128
        M = B * Npad
        convw = [params[4 * i] for i in range(5)] + [params[20]]
        gam = [params[4 * i + 2].detach().float().contiguous() for i in range(5)]
        bet = [params[4 * i + 3].detach().float().contiguous() for i in range(5)]
        cin, cout = convw[0].shape[1], convw[5].shape[0]
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            w0 = convw[0].detach().reshape(64, cin).float().contiguous()
            wsp = [None] + [split_param_cached(lib, convw[i], _CH[i], _CH[i - 1], False, st) for i in range(1, 5)]
            w_last = convw[5].detach().reshape(cout, 256).float().contiguous()
            b_last = params[21].detach().float().contiguous()
            Ys = [torch.empty(M, c, dtype=torch.float32, device=dev) for c in _CH]
            sss = [torch.empty(B, c, 2, dtype=torch.float32, device=dev) for c in _CH]
            mrs = [torch.empty(B, c, 2, dtype=torch.float32, device=dev) for c in _CH]
            stats = torch.zeros(B * 1024 * 2, dtype=torch.float64, device=dev)
            X0 = torch.empty(B, N, cin, dtype=torch.float32, device=dev)
            args, keep = first_layer_args(matches.detach() if matches is not None else None, affine,
                                          [e.detach().float() for e in extras], cin)
            _lib.check(lib.fepe_mlp32_first(*args, w0.data_ptr(), None, Ys[0].data_ptr(), stats.data_ptr(), X0.data_ptr(),
                                            B, N, Npad, 64, st), "fepe_mlp32_first")
            for i in range(5):
                _lib.check(lib.fepe_mlp32_scale_shift(stats.data_ptr(), gam[i].data_ptr(), bet[i].data_ptr(),
                                                      sss[i].data_ptr(), mrs[i].data_ptr(), B, _CH[i], N, 1e-5,
                                                      1 if i < 4 else 0, st), "fepe_mlp32_scale_shift")
                if i < 4:
                    whi, wlo, wsc = wsp[i + 1]
                    _lib.check(lib.fepe_mlp32_gemm(Ys[i].data_ptr(), sss[i].data_ptr(), SLOPE, None, whi.data_ptr(),
                                                   wlo.data_ptr(), wsc.data_ptr(), None, Ys[i + 1].data_ptr(),
                                                   stats.data_ptr(), B, Npad, N, _CH[i], _CH[i + 1], st), "fepe_mlp32_gemm")
            logits = torch.empty(B, cout, N, dtype=torch.float32, device=dev)
            _lib.check(lib.fepe_mlp32_last(Ys[4].data_ptr(), sss[4].data_ptr(), SLOPE, w_last.data_ptr(), b_last.data_ptr(),
                                           logits.data_ptr(), None, B, N, Npad, 256, cout, st), "fepe_mlp32_last")
            del keep
        ctx.dims = (B, N, Npad, cin, cout, affine, n_extras, matches is not None,
                    [1 if e.dim() == 2 else e.shape[2] for e in extras], [tuple(e.shape) for e in extras])
        ctx.saved = (X0, w0, convw, w_last, gam, Ys, sss, mrs)
        ctx.param_objs = params            # the Parameter objects themselves: the backward may add into their .grad
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.lib()
        B, N, Npad, cin, cout, affine, n_extras, has_m, ecs, eshapes = ctx.dims
        X0, w0, convw, w_last, gam, Ys, sss, mrs = ctx.saved
        dev = X0.device
        M = B * Npad
        grads = [None] * 22
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            # every accumulated output of the pass (dW of all layers, conv-bias zeros, A sums, max |dY|) comes out of TWO
            # zero-filled allocations: the step is launch-bound at training batch sizes
            wsz = [64 * cin] + [_CH[i] * _CH[i - 1] for i in range(1, 5)] + [cout * 256, cout]
            fz = torch.zeros(sum(wsz) + 3 * sum(_CH), dtype=torch.float32, device=dev)
            woff = [0]
            for n in wsz:
                woff.append(woff[-1] + n)

            def sink(k):
                """.grad of parameter k when its owner asked for fused accumulation (dist.FlatGradients), else None."""
                prm = ctx.param_objs[k]
                g = prm.grad if getattr(prm, "_fepe_grad_sink", False) else None
                if g is None or g.dtype != torch.float32 or g.device != dev or not g.is_contiguous() or g.data_ptr() % 16:
                    return None
                return g

            widx = [0, 4, 8, 12, 16, 20, 21]                          # parameter index of dWs[0..6]
            wsink = [sink(k) for k in widx]
            dWs = [wsink[i].view(-1) if wsink[i] is not None else fz[woff[i]:woff[i + 1]] for i in range(7)]
            bz, boff = fz[woff[7]:woff[7] + sum(_CH)], 0
            gz = fz[woff[7] + sum(_CH):]                              # dgamma | dbeta of the five blocks (no sink)
            Az = torch.zeros(2 * B * sum(_CH) + 8, dtype=torch.float64, device=dev)
            amax = Az[2 * B * sum(_CH):].view(torch.int32)            # 16 int32 slots (bits of max |dY| per layer)
            aoff = [0]
            for c in _CH:
                aoff.append(aoff[-1] + 2 * B * c)
            dl = dlogits.detach().float().contiguous()
            dX = torch.empty(M, 256, dtype=torch.float32, device=dev)
            dwl, dbl = dWs[5].view(cout, 256), dWs[6]
            _lib.check(lib.fepe_mlp32_last_bwd(dl.data_ptr(), Ys[4].data_ptr(), sss[4].data_ptr(), SLOPE, w_last.data_ptr(),
                                               dX.data_ptr(), dwl.data_ptr(), dbl.data_ptr(), B, N, Npad, 256, cout, st),
                       "fepe_mlp32_last_bwd")
            grads[20] = dwl.reshape(cout, 256, 1) if wsink[5] is None else None
            grads[21] = dbl if wsink[6] is None else None
            dX0 = None
            for i in range(4, -1, -1):
                c = _CH[i]
                A = Az[aoff[i]:aoff[i + 1]].view(B, c, 2)
                dY = torch.empty(M, c, dtype=torch.float32, device=dev)
                am = amax[2 * i:2 * i + 1]
                _lib.check(lib.fepe_mlp32_normbwd(dX.data_ptr(), Ys[i].data_ptr(), sss[i].data_ptr(), mrs[i].data_ptr(),
                                                  gam[i].data_ptr(), SLOPE, A.data_ptr(), dY.data_ptr(), am.data_ptr(),
                                                  B, Npad, N, c, st), "fepe_mlp32_normbwd")
                if sink(4 * i + 1) is None:
                    grads[4 * i + 1] = bz[boff:boff + c]           # conv bias before InstanceNorm: exactly zero gradient
                boff += c
                if i > 0:
                    ci = _CH[i - 1]
                    dW = dWs[i].view(c, ci)
                    _lib.check(lib.fepe_mlp32_wgrad(dY.data_ptr(), am.data_ptr(), Ys[i - 1].data_ptr(),
                                                    sss[i - 1].data_ptr(), SLOPE, dW.data_ptr(), M, Npad, c, ci, st),
                               "fepe_mlp32_wgrad")
                    grads[4 * i] = dW.reshape(c, ci, 1) if wsink[i] is None else None
                    thi, tlo, tsc = split_param_cached(lib, convw[i], c, ci, True, st)   # [ci, c]: the data-gradient "weight"
                    dXn = torch.empty(M, ci, dtype=torch.float32, device=dev)
                    _lib.check(lib.fepe_mlp32_gemm(dY.data_ptr(), None, 1.0, am.data_ptr(), thi.data_ptr(), tlo.data_ptr(),
                                                   tsc.data_ptr(), None, dXn.data_ptr(), None, B, Npad, Npad, c, ci, st),
                               "fepe_mlp32_gemm(dgrad)")
                    dX = dXn
                else:
                    dW = dWs[0].view(64, cin)
                    need_x = any(ctx.needs_input_grad[1:2 + n_extras])
                    dX0 = torch.empty(B, N, cin, dtype=torch.float32, device=dev) if need_x else None
                    _lib.check(lib.fepe_mlp32_first_bwd(dY.data_ptr(), X0.data_ptr(), w0.data_ptr(),
                                                        dX0.data_ptr() if dX0 is not None else None, dW.data_ptr(), B, N,
                                                        Npad, cin, 64, st), "fepe_mlp32_first_bwd")
                    grads[0] = dW.reshape(64, cin, 1) if wsink[0] is None else None
            # dgamma = sum_b A2, dbeta = sum_b A1 for all five blocks: one launch, accumulating into the sinks or into fz
            dgs, dbs, goff = [], [], 0
            for i in range(5):
                sg, sb = sink(4 * i + 2), sink(4 * i + 3)
                dg = sg if sg is not None else gz[goff:goff + _CH[i]]
                db_ = sb if sb is not None else gz[goff + _CH[i]:goff + 2 * _CH[i]]
                goff += 2 * _CH[i]
                dgs.append(dg); dbs.append(db_)
                grads[4 * i + 2] = dg if sg is None else None
                grads[4 * i + 3] = db_ if sb is None else None
            _lib.check(lib.fepe_mlp32_affine_grads(Az.data_ptr(), B, 5, (ctypes.c_int * 5)(*_CH),
                                                   (ctypes.c_void_p * 5)(*[t.data_ptr() for t in dgs]),
                                                   (ctypes.c_void_p * 5)(*[t.data_ptr() for t in dbs]), st),
                       "fepe_mlp32_affine_grads")
        # route the input gradient back to the channel groups
        gm, ge, off = None, [None] * n_extras, 0
        if has_m:
            if ctx.needs_input_grad[1]:
                ax, _, ay, _ = affine                                   # x = ((a m + b) + 1) / 2
                gm = dX0[:, :, 0:4].clone()                             # scalar multiplies: nothing is copied from the
                gm[:, :, 0::2] *= ax / 2                                # host, so the step stays capturable in a graph
                gm[:, :, 1::2] *= ay / 2
            off = 4
        for j in range(n_extras):
            if ctx.needs_input_grad[2 + j]:
                ge[j] = dX0[:, :, off:off + ecs[j]].reshape(eshapes[j])
            off += ecs[j]
        return (None, gm) + tuple(ge) + tuple(grads)


def module_params(fw: nn.Sequential):
    """The 22 parameters of an ErrorEstimator in the order MLP32Function expects."""
    convs, norms = _layers(fw)
    out = []
    for i in range(5):
        out += [convs[i].weight, convs[i].bias, norms[i].weight, norms[i].bias]
    return out + [convs[5].weight, convs[5].bias]


def mlp32_autograd(fw: nn.Sequential, matches, affine, extras, B: int, N: int):
    extras = [e if e.dtype == torch.float32 else e.float() for e in extras]
    return MLP32Function.apply((tuple(float(v) for v in affine) if affine is not None else None, len(extras), B, N),
                               matches, *extras, *module_params(fw))

"""Diagnostics: ops.linear_tc32 forward / dgrad / wgrad against fp64 on GoodCorresNet-shaped layers, and per-parameter
gradient errors of GoodCorresNet (tensor cores vs torch fp32 vs fp64)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200"), os.path.join(ROOT, "tests")]
import torch

from fepe_b200 import ops
from fepe_b200.models import GoodCorresNet

torch.manual_seed(0)
for M, K, Co, sparse in ((1024, 64, 128, False), (1024, 512, 2048, False), (1024, 512, 2048, True), (1024, 4928, 256, False),
                         (1024, 256, 256, False)):
    x = torch.relu(torch.randn(M, K, device="cuda")).requires_grad_(True)
    W = (torch.randn(Co, K, device="cuda") / K ** 0.5).requires_grad_(True)
    gy = torch.randn(M, Co, device="cuda") * 1e-3
    if sparse:
        gy = gy * (torch.rand(M, Co, device="cuda") < 0.002)
    y = ops.linear_tc32(x, W)
    y.backward(gy)
    x64, W64 = x.detach().double(), W.detach().double()
    yr, gxr, gwr = x64 @ W64.t(), gy.double() @ W64, gy.double().t() @ x64
    rel = lambda a, b: float((a.double() - b).norm() / b.norm())
    relm = lambda a, b, mass: float(((a.double() - b).abs() / mass).max())
    # torch fp32 for comparison
    yt = x.detach() @ W.detach().t()
    gxt, gwt = gy @ W.detach(), gy.t() @ x.detach()
    print(f"M={M} K={K} Co={Co} sparse={sparse}: fwd rel {rel(y, yr):.2e} (torch {rel(yt, yr):.2e}); "
          f"dgrad rel {rel(x.grad, gxr):.2e} (torch {rel(gxt, gxr):.2e}); wgrad rel {rel(W.grad, gwr):.2e} (torch {rel(gwt, gwr):.2e})")

import test_goodcorresnet as TG
net = GoodCorresNet(4).cuda().eval()
x = torch.rand(2, 4, 512, device="cuda")
g = torch.randn(2, 1, 512, device="cuda") / 512


def run(module, xin, gout, fn):
    module.zero_grad()
    xx = xin.clone().requires_grad_(True)
    out = fn(module, xx)
    (out * gout).sum().backward()
    return {n: p.grad.double().clone() for n, p in module.named_parameters()}


gp = run(net, x, g, lambda m, t: m(t))
net.use_kernels = False
gp32 = run(net, x, g, lambda m, t: m(t))
net.double()
gpr = run(net, x.double(), g.double(), TG._plain_forward)
for n in gpr:
    nr = float(gpr[n].norm())
    if nr > 1e-12:
        print(f"  {n:28s} |g| {nr:.2e}  tc {float((gp[n] - gpr[n]).norm()) / nr:.2e}  torch32 {float((gp32[n] - gpr[n]).norm()) / nr:.2e}")

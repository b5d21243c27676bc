"""GoodCorresNet (deepFEPE/models/GoodCorresNet.py:35-163; un-instantiable in the reference: its blocks come from the
un-vendored `shaper` package, so PARITY IS UNPINNED -- SURVEY.md 8c).  What can be pinned: the module's surface (ctor
channel spec :45-53, state_dict), its data flow against an independent plain-PyTorch restatement, its tensor-core GEMMs
against fp64, gradients, and the DeepFNet(if_goodCorresArch=True) branch (DeepFNet.py:334-337)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from fepe_b200.models import GoodCorresNet


def _plain_forward(net, x):
    """Independent restatement with stock modules on [B,C,N] tensors (the layout of the reference's forward :95-160)."""
    def block(b, t):
        t = F.conv1d(t, b.conv.weight, b.conv.bias)
        if b.norm is not None:
            t = F.instance_norm(t, weight=b.norm.weight, bias=b.norm.bias, eps=b.norm.eps)
        return F.relu(t) if b.relu else t
    N = x.shape[2]
    feats = []
    for m in net.stem:
        x = block(m, x)
        feats.append(x)
    for m in net.mlp_local:
        x = block(m, x)
        feats.append(x)
    g, _ = torch.max(x, 2, keepdim=True)
    feats.append(g.expand(-1, -1, N))
    x = torch.cat(feats, 1)
    for m in net.mlp_seg:
        x = block(m, x)
    x = block(net.conv_seg, x)
    return F.conv1d(x, net.seg_logit.weight, net.seg_logit.bias)


def test_surface_and_state_dict():
    net = GoodCorresNet(7, bn=False)
    keys = list(net.state_dict().keys())
    assert "stem.0.conv.weight" in keys and "mlp_local.1.norm.weight" in keys and "seg_logit.bias" in keys
    assert net.mlp_seg[0].conv.in_channels == 64 + 128 + 128 + 512 + 2048 + 2048 == 4928     # GoodCorresNet.py:90,139
    assert [m.conv.out_channels for m in net.stem] == [64, 128, 128]                            # :45-53
    assert [m.conv.out_channels for m in net.mlp_local] == [512, 2048]
    assert sum(p.numel() for p in net.parameters()) > 2_000_000
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 7, 16))


@pytest.mark.gpu
@pytest.mark.parametrize("cin,N", [(4, 512), (7, 333)])
def test_forward_and_gradients_against_fp64(cin, N):
    torch.manual_seed(0)
    net = GoodCorresNet(cin).cuda().eval()               # eval: dropout off (deterministic comparison)
    x = torch.rand(2, cin, N, device="cuda")
    g = torch.randn(2, 1, N, device="cuda") / N

    def run(module, xin, gout, fn):
        module.zero_grad()
        xx = xin.clone().requires_grad_(True)
        out = fn(module, xx)
        (out * gout).sum().backward()
        return out.detach().double(), xx.grad.double(), {n: p.grad.double().clone() for n, p in module.named_parameters()}

    out, gx, gp = run(net, x, g, lambda m, t: m(t))                        # tensor-core GEMMs
    net.use_kernels = False
    out32, gx32, gp32 = run(net, x, g, lambda m, t: m(t))                  # same module, PyTorch fp32 matmul
    net.double()
    ref, gxr, gpr = run(net, x.double(), g.double(), _plain_forward)       # independent fp64 restatement
    net.float()
    sc = float(ref.abs().max())
    e_tc, e_32 = float((out - ref).abs().max()) / sc, float((out32 - ref).abs().max()) / sc
    print(f"GoodCorresNet cin={cin} N={N}: logits vs fp64: tensor cores {e_tc:.2e}, torch fp32 {e_32:.2e}")
    assert e_tc < max(5 * e_32, 5e-5)
    worst_tc = max(float((gp[n] - gpr[n]).norm() / gpr[n].norm().clamp_min(1e-30)) for n in gpr if float(gpr[n].norm()) > 1e-12)
    worst_32 = max(float((gp32[n] - gpr[n]).norm() / gpr[n].norm().clamp_min(1e-30)) for n in gpr if float(gpr[n].norm()) > 1e-12)
    print(f"   worst parameter-gradient rel err vs fp64: tensor cores {worst_tc:.2e}, torch fp32 {worst_32:.2e}")
    # The network is not smooth (ReLU kinks, arg-max of the global pooling): one flipped arg-max between an fp32 and the
    # fp64 evaluation moves every upstream gradient by ~3e-3 (measured for BOTH the tensor-core path and torch's fp32
    # matmul, scripts/diag_linear_tc32.py).  The smooth accuracy of the GEMMs is pinned by test_linear_tc32_* below.
    assert worst_tc < 2e-2
    assert float((gx - gxr).norm() / gxr.norm()) < 2e-2


@pytest.mark.gpu
def test_deepfnet_good_corres_arch_branch():
    """DeepFNet(if_goodCorresArch=True) (DeepFNet.py:334-337; the reference builds 6-channel update nets but feeds 7: the
    update net here gets the channel count it receives): forward keys / shapes, rank-2 F at every layer, gradients."""
    from fepe_b200 import synth
    from fepe_b200.models import DeepFNet
    torch.manual_seed(1)
    net = DeepFNet(depth=3, image_size=[376, 1241, 3], if_quality=False, if_goodCorresArch=True).cuda()
    assert isinstance(net.input_weights, GoodCorresNet) and net.update_weights.in_channels == 7
    d = synth.make_batch(2, 256, seed=4)
    outs = net({"matches_xy_ori": torch.from_numpy(d["matches_xy_ori"]).cuda()})
    assert len(outs["out_layers"]) == 3 and outs["weights"].shape == (2, 1, 256)
    for Fo in outs["out_layers"]:
        assert torch.isfinite(Fo).all()
        assert float(torch.linalg.svdvals(Fo.double())[:, 2].max()) < 1e-6
    sum(Fo.pow(2).sum() for Fo in outs["out_layers"]).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


@pytest.mark.gpu
@pytest.mark.parametrize("M,K,Co", [(1024, 64, 128), (1024, 512, 2048), (1024, 4928, 256), (2048, 256, 256)])
def test_linear_tc32_forward_dgrad_wgrad_against_fp64(M, K, Co):
    """ops.linear_tc32 (GoodCorresNet's dense 1x1 convolutions): forward, data gradient and weight gradient on tcgen05
    with split-fp16 operands against fp64 on GoodCorresNet-shaped layers.  The forward bias of long sums (K = 4928:
    1.7e-5) is the truncating fp32 accumulation of tensor memory, 3 K / 16 accumulation steps of ~2^-25 each."""
    from fepe_b200 import ops
    torch.manual_seed(3)
    x = torch.relu(torch.randn(M, K, device="cuda")).requires_grad_(True)
    W = (torch.randn(Co, K, device="cuda") / K ** 0.5).requires_grad_(True)
    gy = torch.randn(M, Co, device="cuda") * 1e-3
    y = ops.linear_tc32(x, W)
    y.backward(gy)
    x64, W64 = x.detach().double(), W.detach().double()
    rel = lambda a, b: float((a.double() - b).norm() / b.norm())
    assert rel(y, x64 @ W64.t()) < 2e-6 + 4e-9 * K              # sums over K
    assert rel(x.grad, gy.double() @ W64) < 2e-6 + 4e-9 * Co      # sums over Co
    assert rel(W.grad, gy.double().t() @ x64) < 2e-6 + 4e-9 * M   # sums over M

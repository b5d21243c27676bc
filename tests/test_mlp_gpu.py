"""Tensor-core MLP (tcgen05 GEMM + InstanceNorm + LeakyReLU + softmax) against the fp32 PyTorch
reference of the same op.  Tolerance: bf16 activations / weights with fp32 accumulation -- logits to
5e-2 absolute on O(1) values, softmax weights to 5 % relative where they matter."""
import numpy as np
import pytest
import torch

from fepe_b200 import _lib, synth
from fepe_b200.models import DeepFNet, ErrorEstimator

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["persist", "persist128", "tile"])
@pytest.mark.parametrize("B,N,K,Co", [(1, 128, 64, 64), (2, 100, 64, 128), (3, 1000, 128, 1024), (2, 1000, 1024, 512),
                                      (2, 333, 512, 256), (40, 1000, 128, 256), (37, 900, 192, 384)])
def test_gemm_against_torch_fp32(B, N, K, Co, mode, dispatch):
    # mode: the persistent kernel (default; 128 x 256 tiles when Co allows), the same with 128 x 128 tiles, and the
    # one-tile-per-CTA kernel (forced through fepe_set_dispatch).  The two larger cases give every
    # persistent CTA several tiles (both TMEM accumulator buffers and the operand ring wrap around).
    dispatch("mlp_gemm", mode)
    lib = _lib.lib()
    torch.manual_seed(0)
    Npad = (N + 127) // 128 * 128
    X = torch.randn(B, Npad, K, device="cuda").bfloat16()
    W = (torch.randn(Co, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(Co, device="cuda")
    Y = torch.full((B * Npad, Co), 7.0, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(B, Co, 2, device="cuda")
    st = lib.fepe_mlp_gemm(X.data_ptr(), W.data_ptr(), bias.data_ptr(), Y.data_ptr(), stats.data_ptr(), B, Npad, N, K, Co,
                           torch.cuda.current_stream().cuda_stream)
    assert st == 0
    torch.cuda.synchronize()
    ref = (X.float().reshape(-1, K) @ W.float().t() + bias).reshape(B, Npad, Co)
    ref[:, N:] = 0
    assert float((Y.float().reshape(B, Npad, Co) - ref).abs().max()) < 0.04       # one bf16 ulp at |y| ~ 8
    refb = ref.bfloat16().float()
    np.testing.assert_allclose(stats[..., 0].cpu(), refb[:, :N].sum(1).cpu(), rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(stats[..., 1].cpu(), (refb[:, :N] ** 2).sum(1).cpu(), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("cin,B,N", [(4, 3, 1000), (7, 2, 333), (4, 1, 128)])
def test_error_estimator_tensor_core_path(cin, B, N):
    torch.manual_seed(1)
    ee = ErrorEstimator(cin).cuda()
    with torch.no_grad():        # non-trivial affine parameters
        for m in ee.fw:
            if isinstance(m, torch.nn.InstanceNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    x = torch.rand(B, cin, N, device="cuda")
    with torch.no_grad():
        ref = ee(x)
        ee.tensor_cores = True
        out = ee(x)
        sm = ee.last_softmax
    assert sm is not None and out.shape == ref.shape
    err = (out - ref).abs()
    print("logits: max abs err %.3e, mean %.3e, ref std %.3f" % (float(err.max()), float(err.mean()), float(ref.std())))
    assert float(err.max()) < 0.1 * max(1.0, float(ref.std())) and float(err.mean()) < 0.02 * max(1.0, float(ref.std()))
    ref_sm = torch.softmax(ref, dim=2)
    np.testing.assert_allclose(sm.sum(2).cpu().numpy(), 1.0, rtol=1e-4)
    big = ref_sm > 0.1 / N
    assert float(((sm - ref_sm).abs() / ref_sm)[big].max()) < 0.15


@pytest.mark.parametrize("variant", ["1", "2"])
@pytest.mark.parametrize("B,N,K,Co", [(2, 100, 64, 128), (3, 1000, 128, 1024), (40, 1000, 1024, 512), (37, 900, 192, 384),
                                      (300, 1000, 64, 128), (150, 1000, 512, 256)])
def test_gemm_norm_fused_equals_norm_then_gemm(B, N, K, Co, variant, dispatch):
    """fepe_mlp_gemm_norm (InstanceNorm + LeakyReLU applied to the operand tiles in shared memory) against
    fepe_mlp_norm followed by fepe_mlp_gemm: the same arithmetic, so Y is bit-identical and the statistics agree to
    fp32 summation order."""
    # variant 1 (default): (a, d) from global memory, 4 transform warps; variant 2: through a shared-memory slot of the
    # stage, 8 transform warps (forced through fepe_set_dispatch)
    dispatch("mlp_fuse", variant)
    lib = _lib.lib()
    torch.manual_seed(2)
    st = torch.cuda.current_stream().cuda_stream
    Npad = (N + 127) // 128 * 128
    Yprev = (torch.randn(B, Npad, K, device="cuda") * 2 + 0.5).bfloat16()
    Yprev[:, N:] = 0
    pstats = torch.stack([Yprev[:, :N].float().sum(1), (Yprev[:, :N].float() ** 2).sum(1)], dim=2).contiguous()
    gamma = torch.empty(K, device="cuda").uniform_(0.5, 1.5)
    beta = torch.empty(K, device="cuda").uniform_(-0.3, 0.3)
    W = (torch.randn(Co, K, device="cuda") / K ** 0.5).bfloat16()
    X = torch.empty(B * Npad, K, device="cuda", dtype=torch.bfloat16)
    Y0 = torch.empty(B * Npad, Co, device="cuda", dtype=torch.bfloat16)
    Y1 = torch.full((B * Npad, Co), 3.0, device="cuda", dtype=torch.bfloat16)
    s0 = torch.zeros(B, Co, 2, device="cuda")
    s1 = torch.zeros(B, Co, 2, device="cuda")
    ss = torch.empty(B, K // 2, 4, device="cuda")
    assert lib.fepe_mlp_norm(Yprev.data_ptr(), pstats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), X.data_ptr(), B, Npad,
                             N, K, 1e-5, 0.01, st) == 0
    assert lib.fepe_mlp_gemm(X.data_ptr(), W.data_ptr(), 0, Y0.data_ptr(), s0.data_ptr(), B, Npad, N, K, Co, st) == 0
    keep = pstats.clone()
    assert lib.fepe_mlp_scale_shift(pstats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ss.data_ptr(), B, K, N, 1e-5, 1,
                                    st) == 0
    assert lib.fepe_mlp_gemm_norm(Yprev.data_ptr(), ss.data_ptr(), 0.01, W.data_ptr(), 0, Y1.data_ptr(), s1.data_ptr(), B,
                                  Npad, N, K, Co, st) == 0
    torch.cuda.synchronize()
    assert float(pstats.abs().max()) == 0.0                      # clear_stats
    mean = keep[..., 0] / N
    a = torch.rsqrt((keep[..., 1] / N - mean * mean).clamp_min(0) + 1e-5) * gamma
    np.testing.assert_allclose(ss.reshape(B, K // 2, 2, 2)[:, :, 0].reshape(B, K).cpu(), a.cpu(), rtol=2e-6)
    np.testing.assert_allclose(ss.reshape(B, K // 2, 2, 2)[:, :, 1].reshape(B, K).cpu(), (beta - mean * a).cpu(), rtol=1e-5,
                               atol=1e-6)
    assert torch.equal(Y0, Y1)
    np.testing.assert_allclose(s1.cpu(), s0.cpu(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("cin,B,N", [(4, 3, 1000), (7, 2, 333), (4, 40, 1000)])
def test_error_estimator_fused_norm_equals_unfused(cin, B, N):
    from fepe_b200.mlp_tc import TensorCoreMLP
    torch.manual_seed(5)
    ee = ErrorEstimator(cin).cuda()
    with torch.no_grad():
        for m in ee.fw:
            if isinstance(m, torch.nn.InstanceNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    x = torch.rand(B, cin, N, device="cuda")
    tc = TensorCoreMLP(ee.fw)
    with torch.no_grad():
        tc.fuse_norm = False
        l0, w0 = tc(x)
        tc.fuse_norm = True
        l1, w1 = tc(x)
    torch.cuda.synchronize()
    # identical bf16 activations layer by layer; only the fp32 statistics are summed in a different order
    assert float((l0 - l1).abs().max()) < 4e-2 * max(1.0, float(l0.std()))     # bf16 + atomics in a different order
    np.testing.assert_allclose(w1.cpu().numpy(), w0.cpu().numpy(), rtol=5e-2, atol=1e-7)


def test_deepfnet_inference_with_tensor_core_mlp():
    torch.manual_seed(3)
    kw = dict(depth=5, image_size=[376, 1241, 3], if_quality=False)
    net = DeepFNet(**kw).cuda()
    d = synth.make_batch(4, 1000, seed=9)
    batch = {"matches_xy_ori": torch.from_numpy(d["matches_xy_ori"]).cuda()}
    with torch.no_grad():
        ref = net(batch)
        net.enable_tensor_core_mlp()
        out = net(batch)
    for l in range(5):
        assert torch.isfinite(out["out_layers"][l]).all()
    # first layer: weights differ by bf16-level noise only, so F agrees to ~1e-2 relative
    from oracle import fepe_oracle as O
    e0 = O.sign_aligned_rel_err(out["out_layers"][0].cpu(), ref["out_layers"][0].cpu())
    print("layer-0 F rel err with bf16 MLP:", e0.tolist())
    assert float(e0.max()) < 5e-2


@pytest.mark.parametrize("M,Co,Ci", [(64, 128, 64), (128, 128, 128), (1024, 1024, 128), (65536, 512, 1024), (4096, 256, 512)])
def test_wgrad_gemm_against_torch(M, Co, Ci):
    """MN-major tcgen05 GEMM: dW = dY^T X with both operands read in place from the row-major activations."""
    lib = _lib.lib()
    torch.manual_seed(0)
    dY = (torch.randn(M, Co, device="cuda") / 8).bfloat16()
    X = torch.randn(M, Ci, device="cuda").bfloat16()
    dW = torch.zeros(Co, Ci, device="cuda")
    assert lib.fepe_mlp_wgrad(dY.data_ptr(), X.data_ptr(), dW.data_ptr(), M, Co, Ci, torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    ref = dY.float().t() @ X.float()
    assert float((dW - ref).abs().max()) < 2e-3 * float(ref.abs().max()) + 1e-3


class _Q(torch.autograd.Function):
    """bf16 storage of a tensor and of its gradient (what the tensor-core path does at every layer boundary)."""

    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _bf16_emulation(ee, x):
    """fp32 PyTorch math with the SAME quantisation points as the tensor-core path: bf16 weights for the four GEMM
    layers, bf16 storage of every pre-norm output Y and block output X' (and of their gradients)."""
    h = x
    convs = [m for m in ee.fw if isinstance(m, torch.nn.Conv1d)]
    norms = [m for m in ee.fw if isinstance(m, torch.nn.InstanceNorm1d)]
    for i in range(5):
        w = convs[i].weight if i == 0 else _Q.apply(convs[i].weight)
        h = _Q.apply(torch.nn.functional.conv1d(h, w, convs[i].bias))
        h = torch.nn.functional.instance_norm(h, weight=norms[i].weight, bias=norms[i].bias, eps=norms[i].eps)
        h = _Q.apply(torch.nn.functional.leaky_relu(h, 0.01))
    return torch.nn.functional.conv1d(h, convs[5].weight, convs[5].bias)


@pytest.mark.parametrize("cin,B,N", [(4, 3, 1000), (7, 2, 333), (4, 2, 256)])
def test_training_path_gradients(cin, B, N):
    """Forward + backward on tensor cores.  Yardstick 1: fp32 PyTorch math with the same bf16 storage points
    (agreement to a few %: same arithmetic up to summation order).  Yardstick 2 (printed): plain fp32 autograd --
    with a RANDOM upstream gradient the parameter gradients are sums with heavy cancellation and bf16 storage
    alone moves them by 10-20 % (the emulation shows the same), so that number is informational."""
    torch.manual_seed(2)
    ee = ErrorEstimator(cin).cuda()
    with torch.no_grad():
        for m in ee.fw:
            if isinstance(m, torch.nn.InstanceNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    x = torch.rand(B, cin, N, device="cuda")
    g = torch.randn(B, 1, N, device="cuda")

    def grads(fn):
        ee.zero_grad()
        xx = x.clone().requires_grad_(True)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            (fn(xx) * g).sum().backward()
        return {n: p.grad.clone() for n, p in ee.named_parameters()}, xx.grad.clone()

    ref32, refx32 = grads(lambda t: ee.fw(t))
    refq, refxq = grads(lambda t: _bf16_emulation(ee, t))
    ee.tensor_cores_training = True
    ours, oursx = grads(lambda t: ee(t))
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-12))
    scale = max(float(v.abs().max()) for v in ref32.values())
    rq, r32 = [], []
    for n in ours:
        if n.endswith("bias") and n not in ("fw.15.bias",) and "fw.%d." % (int(n.split(".")[1])) in ("fw.0.", "fw.3.", "fw.6.", "fw.9.", "fw.12."):
            # conv bias in front of an InstanceNorm: exactly zero gradient (fp32 autograd returns round-off noise)
            assert float(ours[n].abs().max()) == 0.0 and float(ref32[n].abs().max()) < 1e-3 * scale
            continue
        rq.append((n, rel(ours[n], refq[n])))
        r32.append((n, rel(ours[n], ref32[n])))
    print("vs bf16-storage emulation:", "; ".join("%s %.1e" % t for t in rq), "| x %.1e" % rel(oursx, refxq))
    print("vs plain fp32 autograd   :", "; ".join("%s %.1e" % t for t in r32), "| x %.1e" % rel(oursx, refx32))
    assert max(v for _, v in rq) < 1e-1 and rel(oursx, refxq) < 1e-1       # mask flips at |Z| ~ 0 under a random upstream


def test_normbwd_kernel_against_autograd():
    """InstanceNorm(affine) + LeakyReLU adjoint on identical saved tensors: agreement to bf16 output rounding."""
    lib = _lib.lib()
    torch.manual_seed(5)
    B, N, C = 3, 333, 128
    Npad = 384
    Y = torch.zeros(B, Npad, C, device="cuda")
    Y[:, :N] = torch.randn(B, N, C, device="cuda") * 2 + 0.5
    Yb = Y.bfloat16()
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.rand(C, device="cuda") - 0.5
    yv = Yb[:, :N].float().requires_grad_(True)                        # [B,N,C]
    z = torch.nn.functional.instance_norm(yv.permute(0, 2, 1), weight=gamma.clone().requires_grad_(True), bias=beta, eps=1e-5)
    gam_leaf = None
    yv2 = Yb[:, :N].float().permute(0, 2, 1).contiguous().requires_grad_(True)
    g2 = gamma.clone().requires_grad_(True)
    b2 = beta.clone().requires_grad_(True)
    xp = torch.nn.functional.leaky_relu(torch.nn.functional.instance_norm(yv2, weight=g2, bias=b2, eps=1e-5), 0.01)
    dX = torch.randn(B, C, N, device="cuda")
    dXb = torch.zeros(B, Npad, C, device="cuda", dtype=torch.bfloat16)
    dXb[:, :N] = dX.permute(0, 2, 1).bfloat16()
    (xp * dXb[:, :N].float().permute(0, 2, 1)).sum().backward()
    Xp = torch.zeros(B, Npad, C, device="cuda", dtype=torch.bfloat16)
    Xp[:, :N] = xp.detach().permute(0, 2, 1).bfloat16()
    stats = torch.stack((Yb[:, :N].float().sum(1), (Yb[:, :N].float() ** 2).sum(1)), 2).contiguous()
    A = torch.zeros(B, C, 2, device="cuda")
    dY = torch.empty(B, Npad, C, device="cuda", dtype=torch.bfloat16)
    assert lib.fepe_mlp_normbwd(dXb.data_ptr(), Xp.data_ptr(), Yb.data_ptr(), stats.data_ptr(), gamma.data_ptr(), A.data_ptr(),
                                dY.data_ptr(), B, Npad, N, C, 1e-5, 0.01, torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    ref_dy = yv2.grad.permute(0, 2, 1)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(dY[:, :N].float(), ref_dy) < 1e-2
    assert float(dY[:, N:].float().abs().max()) == 0.0
    assert rel(A[:, :, 1].sum(0), g2.grad) < 2e-3 and rel(A[:, :, 0].sum(0), b2.grad) < 2e-3


def test_first_and_last_layer_backward_kernels():
    lib = _lib.lib()
    torch.manual_seed(6)
    st = torch.cuda.current_stream().cuda_stream
    B, N, Npad = 2, 300, 384
    # last layer
    X = torch.zeros(B, Npad, 256, device="cuda", dtype=torch.bfloat16)
    X[:, :N] = torch.randn(B, N, 256, device="cuda").bfloat16()
    w = torch.randn(256, device="cuda") / 16
    dl = torch.randn(B, N, device="cuda")
    dX = torch.empty(B, Npad, 256, device="cuda", dtype=torch.bfloat16)
    dw, db = torch.zeros(256, device="cuda"), torch.zeros(1, device="cuda")
    assert lib.fepe_mlp_last_bwd(dl.data_ptr(), X.data_ptr(), w.data_ptr(), dX.data_ptr(), dw.data_ptr(), db.data_ptr(), B, N,
                                 Npad, 256, st) == 0
    torch.cuda.synchronize()
    np.testing.assert_allclose(dw.cpu(), torch.einsum("bn,bnk->k", dl, X[:, :N].float()).cpu(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(float(db), float(dl.sum()), rtol=1e-5, atol=1e-4)
    ref = (dl[:, :, None] * w).bfloat16().float()
    assert float((dX[:, :N].float() - ref).abs().max()) < 1e-6 and float(dX[:, N:].float().abs().max()) == 0.0
    # first layer
    Ci = 7
    dY = torch.zeros(B, Npad, 64, device="cuda", dtype=torch.bfloat16)
    dY[:, :N] = torch.randn(B, N, 64, device="cuda").bfloat16()
    X0 = torch.rand(B, N, Ci, device="cuda")
    W = torch.randn(64, Ci, device="cuda")
    dX0 = torch.zeros(B, N, Ci, device="cuda")
    dW = torch.zeros(64, Ci, device="cuda")
    assert lib.fepe_mlp_first_bwd(dY.data_ptr(), X0.data_ptr(), W.data_ptr(), dX0.data_ptr(), dW.data_ptr(), B, N, Npad, Ci, 64,
                                  st) == 0
    torch.cuda.synchronize()
    np.testing.assert_allclose(dX0.cpu(), (dY[:, :N].float() @ W).cpu(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dW.cpu(), torch.einsum("bnc,bnk->ck", dY[:, :N].float(), X0).cpu(), rtol=1e-4, atol=1e-3)


def test_training_path_smooth_upstream_against_fp32():
    """With a non-cancelling upstream gradient (d/dlogits of logsumexp = softmax > 0) bf16 storage costs a few
    per cent at most against plain fp32 autograd."""
    torch.manual_seed(4)
    ee = ErrorEstimator(4).cuda()
    x = torch.rand(3, 4, 1000, device="cuda")

    def grads():
        ee.zero_grad()
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            torch.logsumexp(ee(x) * 3.0, dim=2).sum().backward()
        return {n: p.grad.clone() for n, p in ee.named_parameters()}
    ref = grads()
    ee.tensor_cores_training = True
    ours = grads()
    scale = max(float(v.abs().max()) for v in ref.values())
    worst = 0.0
    for n in ref:
        if float(ref[n].abs().max()) < 1e-3 * scale:
            continue
        worst = max(worst, float((ours[n] - ref[n]).norm() / ref[n].norm()))
    print("smooth upstream: worst parameter-gradient rel err vs fp32 %.3e" % worst)
    assert worst < 6e-2

"""GPU tests of the fp32-parity tensor-core MLP (csrc/fepe_mlp32.cu, fepe_b200/mlp32.py): every kernel against an fp64
evaluation of the same arithmetic, the whole ErrorEstimator against fp64 AND against PyTorch's fp32 kernels (the
reference's arithmetic) -- the split-fp16 path must be in the accuracy class of fp32, not of bf16."""
import numpy as np
import pytest
import torch

from fepe_b200 import _lib, mlp32, ops
from fepe_b200.models import ErrorEstimator

pytestmark = pytest.mark.gpu
SLOPE = 0.01


def _st():
    return torch.cuda.current_stream().cuda_stream


def test_prepare_weights_splits_to_22_bits():
    lib = _lib.lib()
    torch.manual_seed(0)
    for Co, K, scale in ((128, 64, 0.1), (1024, 128, 3e-3), (512, 1024, 40.0), (4, 7, 1.0)):
        W = (torch.randn(Co, K, device="cuda") * scale)
        W[0, 0] = 0.0
        whi, wlo, wsc = mlp32.split_weight(lib, W, _st())
        torch.cuda.synchronize()
        s, inv = float(wsc[0]), float(wsc[1])
        assert s * inv == 1.0 and np.log2(s) == round(np.log2(s))           # an exact power of two
        amax = float(W.abs().max()) * s
        assert 2 ** 13 <= amax < 2 ** 14, amax
        rec = (whi.double() + wlo.double()) * inv
        # hi + lo keeps 22 bits of every weight; below 2^-14 / s the lo part is an fp16 subnormal (absolute floor 2^-25 / s,
        # i.e. 2^-38 of the largest weight)
        err = (rec - W.double()).abs() - (2.0 ** -21 * W.double().abs() + 2.0 ** -24 * inv)
        assert float(err.max()) <= 0.0, float(err.max())
        assert float(rec[0, 0]) == 0.0
        assert torch.isfinite(whi.float()).all() and torch.isfinite(wlo.float()).all()


def _ref_gemm(Yprev, ss, W, bias, B, Npad, N):
    """fp64: X' = LeakyReLU(a y + d) (or y), Y = X' W^T + b, padded rows zero; stats over the N real rows."""
    y = Yprev.double().reshape(B, Npad, -1)
    if ss is not None:
        a, d = ss.double()[:, :, 0].unsqueeze(1), ss.double()[:, :, 1].unsqueeze(1)
        t = y * a + d
        y = torch.maximum(t, SLOPE * t)
    out = y @ W.double().t()
    if bias is not None:
        out = out + bias.double()
    out[:, N:] = 0
    stats = torch.stack((out[:, :N].sum(1), (out[:, :N] ** 2).sum(1)), 2)
    return out.reshape(B * Npad, -1), stats


@pytest.mark.parametrize("B,N,K,Co,mode", [
    (1, 128, 64, 64, "norm"), (2, 100, 64, 128, "norm"), (3, 1000, 128, 1024, "norm"), (2, 1000, 1024, 512, "norm"),
    (2, 333, 512, 256, "norm"), (40, 1000, 128, 256, "norm"), (37, 900, 192, 384, "bias"), (5, 700, 256, 192, "norm"),
    (160, 1000, 64, 128, "norm"), (3, 1000, 128, 64, "identity"), (2, 640, 1024, 128, "identity"),
    (2, 512, 4928, 256, "norm"),
])
def test_gemm_matches_fp64(B, N, K, Co, mode):
    """Y = act(Yprev) W^T on tensor cores with split operands vs the fp64 result: error a few 2^-22 of the row's
    |x'| . |w| mass (what an fp32 FMA chain gives), statistics to fp32 round-off of their magnitude.  The larger cases
    give every persistent CTA several tiles (both TMEM accumulator buffers and the operand ring wrap around); 4928 is
    GoodCorresNet's concatenated width."""
    lib = _lib.lib()
    torch.manual_seed(1)
    Npad = (N + 127) // 128 * 128
    Yprev = torch.randn(B, Npad, K, device="cuda") * 2 + 0.5
    Yprev[:, N:] = 0
    W = torch.randn(Co, K, device="cuda") / K ** 0.5
    ss = bias = None
    if mode != "identity":
        ss = torch.stack((torch.rand(B, K, device="cuda") + 0.5, torch.randn(B, K, device="cuda")), 2).contiguous()
    if mode == "bias":
        bias = torch.randn(Co, device="cuda")
    nvalid = Npad if mode == "identity" else N
    whi, wlo, wsc = mlp32.split_weight(lib, W, _st())
    Y = torch.full((B * Npad, Co), 7.0, device="cuda")
    stats = torch.zeros(B, Co, 2, device="cuda", dtype=torch.float64)
    st = lib.fepe_mlp32_gemm(Yprev.data_ptr(), ss.data_ptr() if ss is not None else None, SLOPE, None, whi.data_ptr(),
                             wlo.data_ptr(), wsc.data_ptr(), bias.data_ptr() if bias is not None else None, Y.data_ptr(),
                             stats.data_ptr(), B, Npad, nvalid, K, Co, _st())
    assert st == 0, st
    torch.cuda.synchronize()
    ref, rstats = _ref_gemm(Yprev, ss, W, bias, B, Npad, nvalid)
    # per-row error yardstick: sum_k |x'_k| |w_ck|
    y = Yprev.double().reshape(B, Npad, K)
    if ss is not None:
        t = y * ss.double()[:, :, 0].unsqueeze(1) + ss.double()[:, :, 1].unsqueeze(1)
        y = torch.maximum(t, SLOPE * t)
    mass = (y.abs() @ W.double().abs().t()).reshape(B * Npad, Co) + 1e-30
    rel = ((Y.double() - ref).abs() / mass)
    rel[ref == 0] = 0
    print(f"B={B} N={N} K={K} Co={Co} {mode}: max err / mass = {float(rel.max()):.2e} (2^-22 = 2.4e-7)")
    # measured 4e-7 (K = 64) .. 1.2e-6 (K = 1024) .. 2.8e-6 (K = 4928): product errors 3 x 2^-22 plus the fp32 accumulation
    # of 3 K / 16 partial MMAs in tensor memory
    assert float(rel.max()) < 1.0e-6 + 5e-10 * K, float(rel.max())
    if nvalid < Npad:
        assert float(Y.reshape(B, Npad, Co)[:, N:].abs().max()) == 0.0
    serr = (stats - rstats).abs() / (rstats.abs() + 1e-300)
    # sums: fp32 partial sums over 128 rows folded in fp64
    # the statistics are taken of the fp32 values the kernel stored: exact (fp64) sums of THOSE
    yk = Y.double().reshape(B, Npad, Co)[:, :nvalid]
    kstats = torch.stack((yk.sum(1), (yk ** 2).sum(1)), 2)
    # (fp32 partial sums over the 128 rows of a tile, pivoted; folded in fp64: measured 2.7e-7 .. 4.8e-7)
    assert float(((stats[..., 0] - kstats[..., 0]).abs() / (yk.abs().sum(1) + 1e-30)).max()) < 3e-6
    assert float(((stats[..., 1] - kstats[..., 1]).abs() / (kstats[..., 1] + 1e-300)).max()) < 3e-6
    # against the fp64 result: dominated by the tensor core's truncating fp32 accumulation (3 K / 16 steps, bias ~2^-25 each)
    assert float(serr[..., 1].max()) < 1e-5 + 2e-8 * K


def test_statistics_survive_a_large_mean():
    """|mean| >> std: E[y^2] - mean^2 loses everything in fp32; the pivoted fp64 statistics keep the variance (the
    resulting (a, d) reproduce torch's InstanceNorm, which uses a two-pass variance)."""
    lib = _lib.lib()
    torch.manual_seed(2)
    B, N, K, Co = 2, 1000, 64, 128
    Npad = 1024
    Yprev = torch.randn(B, Npad, K, device="cuda") * 1e-3 + 5.0       # post-activation mean 5, std 1e-3
    Yprev[:, N:] = 0
    W = torch.randn(Co, K, device="cuda") / K ** 0.5
    whi, wlo, wsc = mlp32.split_weight(lib, W, _st())
    Y = torch.empty(B * Npad, Co, device="cuda")
    stats = torch.zeros(B, Co, 2, device="cuda", dtype=torch.float64)
    assert lib.fepe_mlp32_gemm(Yprev.data_ptr(), None, 1.0, None, whi.data_ptr(), wlo.data_ptr(), wsc.data_ptr(), None,
                               Y.data_ptr(), stats.data_ptr(), B, Npad, N, K, Co, _st()) == 0
    gamma, beta = torch.rand(Co, device="cuda") + 0.5, torch.randn(Co, device="cuda")
    ss = torch.empty(B, Co, 2, device="cuda")
    assert lib.fepe_mlp32_scale_shift(stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ss.data_ptr(), None, B, Co, N,
                                      1e-5, 0, _st()) == 0
    torch.cuda.synchronize()
    y = Y.reshape(B, Npad, Co)[:, :N].double()                       # the values the statistics were taken of
    mean, var = y.mean(1), y.var(1, unbiased=False)
    a = gamma.double() / torch.sqrt(var + 1e-5)
    d = beta.double() - mean * a
    assert float(((ss[..., 0].double() - a).abs() / a.abs()).max()) < 1e-5
    # d = beta - mean a is large (mean / std ~ 1e3): compare the normalised VALUES, which is what matters downstream
    xn = y * ss[..., 0].double().unsqueeze(1) + ss[..., 1].double().unsqueeze(1)
    xr = y * a.unsqueeze(1) + d.unsqueeze(1)
    assert float((xn - xr).abs().max()) < 2e-3 * float(xr.abs().max())


@pytest.mark.parametrize("N", [1000, 37])
def test_first_layer_reads_the_model_inputs_in_place(N):
    lib = _lib.lib()
    torch.manual_seed(3)
    B, Npad = 3, (N + 127) // 128 * 128
    m = torch.rand(B, N, 4, device="cuda") * torch.tensor([1241., 376., 1241., 376.], device="cuda")
    aff = ops.hw_affine([376, 1241])
    q = torch.rand(B, N, 2, device="cuda")
    w, e, r = torch.rand(B, N, device="cuda") / N, torch.rand(B, N, device="cuda") * 0.5, torch.randn(B, N, device="cuda") * 1e-3
    W = torch.randn(64, 9, device="cuda")
    Y = torch.empty(B * Npad, 64, device="cuda")
    stats = torch.zeros(B, 64, 2, device="cuda", dtype=torch.float64)
    args, keep = mlp32.first_layer_args(m, aff, [q, w, e, r], 9)
    X0 = torch.empty(B, N, 9, device="cuda")
    assert lib.fepe_mlp32_first(*args, W.data_ptr(), None, Y.data_ptr(), stats.data_ptr(), X0.data_ptr(), B, N, Npad, 64,
                                _st()) == 0
    torch.cuda.synchronize()
    feat = ErrorEstimator._features(m, aff, [q, w, e, r]).double()                   # [B,9,N], the reference's cat
    assert float((X0.double() - feat.permute(0, 2, 1)).abs().max()) < 1e-6
    ref = torch.einsum("oc,bcn->bno", W.double(), feat)
    got = Y.reshape(B, Npad, 64)
    assert float((got[:, :N].double() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
    assert float(got[:, N:].abs().max() if Npad > N else 0.0) == 0.0
    assert torch.allclose(stats[..., 0], ref.sum(1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(stats[..., 1], (ref ** 2).sum(1), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("Co", [1, 4])
def test_last_layer_and_softmax(Co):
    lib = _lib.lib()
    torch.manual_seed(4)
    B, N, Npad = 3, 900, 1024
    Y = torch.randn(B, Npad, 256, device="cuda")
    ss = torch.stack((torch.rand(B, 256, device="cuda") + 0.5, torch.randn(B, 256, device="cuda")), 2).contiguous()
    W, bias = torch.randn(Co, 256, device="cuda") / 16, torch.randn(Co, device="cuda")
    logits = torch.empty(B, Co, N, device="cuda")
    weights = torch.empty(B, 1, N, device="cuda") if Co == 1 else None
    assert lib.fepe_mlp32_last(Y.data_ptr(), ss.data_ptr(), SLOPE, W.data_ptr(), bias.data_ptr(), logits.data_ptr(),
                               weights.data_ptr() if weights is not None else None, B, N, Npad, 256, Co, _st()) == 0
    torch.cuda.synchronize()
    t = Y.double()[:, :N] * ss.double()[:, :, 0].unsqueeze(1) + ss.double()[:, :, 1].unsqueeze(1)
    x = torch.maximum(t, SLOPE * t)
    ref = (x @ W.double().t() + bias.double()).permute(0, 2, 1)
    assert float((logits.double() - ref).abs().max()) < 1e-5
    if Co == 1:
        assert float((weights.double() - torch.softmax(ref, 2)).abs().max() / torch.softmax(ref, 2).max()) < 1e-5


@pytest.mark.parametrize("cin,cout,B,N", [(4, 1, 3, 1000), (7, 1, 2, 1000), (9, 1, 2, 333), (7, 4, 2, 500), (4, 1, 64, 1000)])
def test_error_estimator_is_in_the_fp32_accuracy_class(cin, cout, B, N):
    """The default path (tc32) vs fp64 truth, next to PyTorch's own fp32 kernels vs the same truth: the logits of the
    split-fp16 tensor-core path must be as close to fp64 as fp32 arithmetic is (same order of magnitude), i.e. ~1e-6
    relative -- three orders below the bf16 path's 3e-2."""
    torch.manual_seed(5)
    ee = ErrorEstimator(cin, cout).cuda()
    with torch.no_grad():
        for m in ee.fw:
            if isinstance(m, torch.nn.InstanceNorm1d):                   # non-trivial affine, as after training
                m.weight.copy_(torch.rand_like(m.weight) + 0.5)
                m.bias.copy_(torch.randn_like(m.bias) * 0.3)
    x = torch.rand(B, cin, N, device="cuda")
    x[:, 4:] = x[:, 4:] * 1e-2                                            # weights / residual channels are small numbers
    with torch.no_grad():
        assert ee.path == "tc32"
        got = ee(x)
        sm = ee.last_softmax
        ee.set_path("torch")
        lib32 = ee(x)
        truth = ee.double()(x.double())              # fp64 through PyTorch's kernels (path "torch")
        ee.float().set_path("tc32")
    scale = float(truth.abs().max())
    e_tc, e_lib = float((got.double() - truth).abs().max()) / scale, float((lib32.double() - truth).abs().max()) / scale
    print(f"cin={cin} cout={cout} B={B} N={N}: |logits| {scale:.2f}; tc32 vs fp64 {e_tc:.2e}, torch fp32 vs fp64 {e_lib:.2e}")
    assert e_tc < max(4 * e_lib, 2e-5), (e_tc, e_lib)
    assert e_tc < 1e-4
    if cout == 1:
        ref_sm = torch.softmax(truth, 2)
        assert float(((sm.double() - ref_sm).abs() / ref_sm).max()) < max(20 * e_tc * scale, 1e-4)


# ------------------------------------------------------------------------------------------------ backward
def test_normbwd_matches_fp64_autograd():
    lib = _lib.lib()
    torch.manual_seed(6)
    for B, N, C in ((2, 1000, 128), (3, 333, 1024), (2, 128, 64), (2, 900, 512)):
        Npad = (N + 127) // 128 * 128
        y = torch.randn(B, Npad, C, device="cuda") * 2 + 0.3
        y[:, N:] = 0
        gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.3
        dX = torch.randn(B, Npad, C, device="cuda") * 1e-4                  # gradients are small numbers
        dX[:, N:] = 0
        # forward statistics through the product's own kernel
        yv = y[:, :N].double()
        stats = torch.stack((yv.sum(1), (yv ** 2).sum(1)), 2).contiguous()
        ss, mr = torch.empty(B, C, 2, device="cuda"), torch.empty(B, C, 2, device="cuda")
        assert lib.fepe_mlp32_scale_shift(stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ss.data_ptr(), mr.data_ptr(),
                                          B, C, N, 1e-5, 0, _st()) == 0
        A = torch.zeros(B, C, 2, device="cuda", dtype=torch.float64)
        dY = torch.full((B, Npad, C), 7.0, device="cuda")
        amax = torch.zeros(1, dtype=torch.int32, device="cuda")
        assert lib.fepe_mlp32_normbwd(dX.data_ptr(), y.data_ptr(), ss.data_ptr(), mr.data_ptr(), gamma.data_ptr(), SLOPE,
                                      A.data_ptr(), dY.data_ptr(), amax.data_ptr(), B, Npad, N, C, _st()) == 0
        torch.cuda.synchronize()
        y64 = y[:, :N].double().permute(0, 2, 1).clone().requires_grad_(True)          # [B,C,N]
        g64, b64 = gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
        out = torch.nn.functional.leaky_relu(torch.nn.functional.instance_norm(y64, weight=g64, bias=b64, eps=1e-5), SLOPE)
        out.backward(dX[:, :N].double().permute(0, 2, 1))
        ref = y64.grad.permute(0, 2, 1)
        sc = float(ref.abs().max())
        assert float((dY[:, :N].double() - ref).abs().max()) < 2e-5 * sc, (B, N, C)
        assert float(dY[:, N:].abs().max() if Npad > N else 0.0) == 0.0
        assert float((A[:, :, 1].sum(0) - g64.grad).abs().max()) < 1e-5 * float(g64.grad.abs().max())
        assert float((A[:, :, 0].sum(0) - b64.grad).abs().max()) < 1e-5 * float(b64.grad.abs().max())
        got_max = float(amax.view(torch.float32)[0])
        assert abs(got_max - float(dY.abs().max())) <= 1e-6 * got_max


@pytest.mark.parametrize("tile", ["128", "256"])
@pytest.mark.parametrize("B,N,Co,Ci", [(1, 128, 128, 64), (2, 1000, 1024, 128), (64, 1000, 512, 1024), (4, 1000, 256, 512),
                                       (2, 300, 128, 64), (16, 1000, 128, 256), (1, 128, 256, 768)])
def test_wgrad_matches_fp64(B, N, Co, Ci, tile, dispatch):
    """dW = dY^T LeakyReLU(a Yprev + d) on tcgen05 with both operands MN-major, split into fp16 pairs in place; 128 x 128
    / 128 x 64 output tiles and, where Ci % 256 == 0, 128 x 256 (the automatic choice)."""
    if tile == "256" and Ci % 256:
        pytest.skip("128 x 256 tiles need Ci % 256 == 0")
    dispatch("wgrad", tile)
    lib = _lib.lib()
    torch.manual_seed(7)
    Npad = (N + 127) // 128 * 128
    M = B * Npad
    dY = torch.randn(B, Npad, Co, device="cuda") * 3e-6                    # far below fp16's normal range: needs the scale
    dY[:, N:] = 0
    Yp = torch.randn(B, Npad, Ci, device="cuda") * 2 + 0.5
    ss = torch.stack((torch.rand(B, Ci, device="cuda") + 0.5, torch.randn(B, Ci, device="cuda")), 2).contiguous()
    amax = dY.abs().max().reshape(1).view(torch.int32).clone()
    dW = torch.zeros(Co, Ci, device="cuda")
    assert lib.fepe_mlp32_wgrad(dY.data_ptr(), amax.data_ptr(), Yp.data_ptr(), ss.data_ptr(), SLOPE, dW.data_ptr(), M, Npad,
                                Co, Ci, _st()) == 0
    torch.cuda.synchronize()
    t = Yp.double() * ss.double()[:, :, 0].unsqueeze(1) + ss.double()[:, :, 1].unsqueeze(1)
    x = torch.maximum(t, SLOPE * t).reshape(M, Ci)
    ref = dY.double().reshape(M, Co).t() @ x
    mass = dY.double().abs().reshape(M, Co).t() @ x.abs()
    rel = float(((dW.double() - ref).abs() / mass).max())
    print(f"wgrad B={B} N={N} {Co}x{Ci}: max err / mass {rel:.2e}")
    assert rel < 2e-6, rel


def test_dgrad_gemm_with_tiny_operand():
    """Data gradient dX = dY W through fepe_mlp32_gemm (ss = NULL) with dY ~ 1e-7: the power-of-two pre-scale from
    a_amax keeps the fp16 split exact."""
    lib = _lib.lib()
    torch.manual_seed(8)
    B, N, K, Co = 3, 1000, 512, 1024
    Npad = 1024
    dY = torch.randn(B, Npad, K, device="cuda") * 1e-7
    W = torch.randn(Co, K, device="cuda") / K ** 0.5
    whi, wlo, wsc = mlp32.split_weight(lib, W, _st())
    amax = dY.abs().max().reshape(1).view(torch.int32).clone()
    Y = torch.empty(B * Npad, Co, device="cuda")
    assert lib.fepe_mlp32_gemm(dY.data_ptr(), None, 1.0, amax.data_ptr(), whi.data_ptr(), wlo.data_ptr(), wsc.data_ptr(), None,
                               Y.data_ptr(), None, B, Npad, Npad, K, Co, _st()) == 0
    torch.cuda.synchronize()
    ref = dY.double().reshape(-1, K) @ W.double().t()
    mass = dY.double().abs().reshape(-1, K) @ W.double().abs().t()
    assert float(((Y.double() - ref).abs() / mass).max()) < 1.5e-6


@pytest.mark.parametrize("cin,cout,B,N", [(4, 1, 3, 1000), (7, 1, 2, 777), (9, 4, 2, 500)])
def test_error_estimator_gradients_are_in_the_fp32_accuracy_class(cin, cout, B, N):
    """Training path (tc32 forward + backward kernels, no library kernel) vs fp64 autograd, next to PyTorch's fp32
    autograd vs the same truth: parameter and input gradients."""
    torch.manual_seed(9)
    ee = ErrorEstimator(cin, cout).cuda()
    with torch.no_grad():
        for m in ee.fw:
            if isinstance(m, torch.nn.InstanceNorm1d):
                m.weight.copy_(torch.rand_like(m.weight) + 0.5)
                m.bias.copy_(torch.randn_like(m.bias) * 0.3)
    x = torch.rand(B, cin, N, device="cuda")
    g = torch.randn(B, cout, N, device="cuda") / N

    def grads(module, xin, gout):
        module.zero_grad()
        xx = xin.clone().requires_grad_(True)
        out = module(xx)
        (out * gout).sum().backward()
        return {n: p.grad.detach().double().clone() for n, p in module.named_parameters()}, xx.grad.detach().double(), out.detach().double()

    assert ee.path == "tc32"
    ours, oursx, out = grads(ee, x, g)
    ee.set_path("torch")
    lib32, lib32x, out32 = grads(ee, x, g)
    ee.double()
    truth, truthx, out64 = grads(ee, x.double(), g.double())
    ee.float().set_path("tc32")
    assert float((out - out64).abs().max()) < 1e-4 * float(out64.abs().max())
    worst_o, worst_l = 0.0, 0.0
    scale = max(float(v.abs().max()) for v in truth.values())
    for n, ref in truth.items():
        if float(ref.abs().max()) < 1e-6 * scale:            # conv biases in front of an InstanceNorm: zero gradient
            assert float(ours[n].abs().max()) <= 1e-5 * scale, n
            continue
        eo = float((ours[n] - ref).norm() / ref.norm())
        el = float((lib32[n] - ref).norm() / ref.norm())
        worst_o, worst_l = max(worst_o, eo), max(worst_l, el)
    ex_o = float((oursx - truthx).norm() / truthx.norm())
    ex_l = float((lib32x - truthx).norm() / truthx.norm())
    print(f"cin={cin} cout={cout}: worst parameter-gradient rel err vs fp64: tc32 {worst_o:.2e}, torch fp32 {worst_l:.2e}; "
          f"input gradient: tc32 {ex_o:.2e}, torch fp32 {ex_l:.2e}")
    assert worst_o < max(5 * worst_l, 2e-4), (worst_o, worst_l)
    assert ex_o < max(5 * ex_l, 2e-4), (ex_o, ex_l)


def test_deepfnet_training_step_on_the_kernel_path():
    """A DeepFNet training step (depth 3, quality channel) with the default tc32 path: no parameter is left without a
    gradient, and the gradients agree with the same step through PyTorch's fp32 kernels."""
    from fepe_b200 import synth
    from fepe_b200.models import DeepFNet
    from oracle import fepe_oracle as O
    torch.manual_seed(10)
    net = DeepFNet(depth=3, image_size=[376, 1241, 3], if_quality=True, quality_size=1).cuda()
    d = synth.make_batch(4, 640, seed=21)
    batch = {"matches_xy_ori": torch.from_numpy(d["matches_xy_ori"]).cuda(), "quality": torch.rand(4, 640, 1, device="cuda")}
    v1, v2 = torch.from_numpy(d["pts1_virt"]).cuda(), torch.from_numpy(d["pts2_virt"]).cuda()

    def step():
        net.zero_grad()
        outs = net(batch)
        p1 = (outs["T1"] @ v1.permute(0, 2, 1)).permute(0, 2, 1)
        p2 = (outs["T1"] @ v2.permute(0, 2, 1)).permute(0, 2, 1)
        loss = sum(O.epi_residual(p1, p2, Fo, 0.02).mean() for Fo in outs["out_layers"]) / 3
        loss.backward()
        return float(loss), {n: p.grad.detach().clone() for n, p in net.named_parameters()}

    l_tc, g_tc = step()
    net.set_mlp_path("torch")
    l_lib, g_lib = step()
    assert abs(l_tc - l_lib) < 1e-4 * abs(l_lib)
    scale = max(float(v.abs().max()) for v in g_lib.values())
    for n, ref in g_lib.items():
        assert torch.isfinite(g_tc[n]).all(), n
        if float(ref.abs().max()) < 1e-4 * scale:
            continue
        rel = float((g_tc[n] - ref).norm() / ref.norm())
        assert rel < 2e-2, (n, rel)          # two fp32 evaluations of an ill-conditioned recursion; fp64 is the judge above


def test_weight_split_cache_does_not_survive_the_parameter():
    """The per-parameter cache of split weights is keyed by id(): a NEW model whose parameters land on the addresses of a
    dead one (same id, same data pointer, same version) must not be served the old model's weights."""
    import gc
    x = torch.rand(2, 4, 256, device="cuda", requires_grad=True)
    outs = []
    for seed in (1, 2, 3, 4):
        torch.manual_seed(seed)
        ee = ErrorEstimator(4).cuda()
        y_train = ee(x)                                   # autograd path: fills the cache
        with torch.no_grad():
            y_inf = ee(x)                                 # inference path: its own per-instance copies
        assert float((y_train - y_inf).abs().max()) < 1e-5 * float(y_inf.abs().max()), seed
        outs.append(y_inf)
        del ee, y_train
        gc.collect()
    assert float((outs[0] - outs[1]).abs().max()) > 1e-3      # different seeds do give different networks

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
        err = ((rec - W.double()).abs() / W.double().abs().clamp_min(1e-30))[W != 0]
        assert float(err.max()) < 2.0 ** -21, float(err.max())
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
    st = lib.fepe_mlp32_gemm(Yprev.data_ptr(), ss.data_ptr() if ss is not None else None, SLOPE, whi.data_ptr(),
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
    assert float(rel.max()) < 1.5e-6, float(rel.max())
    if nvalid < Npad:
        assert float(Y.reshape(B, Npad, Co)[:, N:].abs().max()) == 0.0
    serr = (stats - rstats).abs() / (rstats.abs() + 1e-300)
    # sums: fp32 partial sums over 128 rows folded in fp64
    assert float(((stats[..., 0] - rstats[..., 0]).abs() / (ref.reshape(B, Npad, Co).abs().sum(1) + 1e-30)).max()) < 1e-6
    assert float(serr[..., 1].max()) < 1e-6


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
    assert lib.fepe_mlp32_gemm(Yprev.data_ptr(), None, 1.0, whi.data_ptr(), wlo.data_ptr(), wsc.data_ptr(), None,
                               Y.data_ptr(), stats.data_ptr(), B, Npad, N, K, Co, _st()) == 0
    gamma, beta = torch.rand(Co, device="cuda") + 0.5, torch.randn(Co, device="cuda")
    ss = torch.empty(B, Co, 2, device="cuda")
    assert lib.fepe_mlp32_scale_shift(stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ss.data_ptr(), B, Co, N, 1e-5,
                                      0, _st()) == 0
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
    assert lib.fepe_mlp32_first(*args, W.data_ptr(), None, Y.data_ptr(), stats.data_ptr(), B, N, Npad, 64, _st()) == 0
    torch.cuda.synchronize()
    feat = ErrorEstimator._features(m, aff, [q, w, e, r]).double()                   # [B,9,N], the reference's cat
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
        ee.set_path("tc32")
        truth = ee.double()(x.double())
        ee.float()
    scale = float(truth.abs().max())
    e_tc, e_lib = float((got.double() - truth).abs().max()) / scale, float((lib32.double() - truth).abs().max()) / scale
    print(f"cin={cin} cout={cout} B={B} N={N}: |logits| {scale:.2f}; tc32 vs fp64 {e_tc:.2e}, torch fp32 vs fp64 {e_lib:.2e}")
    assert e_tc < max(4 * e_lib, 2e-5), (e_tc, e_lib)
    assert e_tc < 1e-4
    if cout == 1:
        ref_sm = torch.softmax(truth, 2)
        assert float(((sm.double() - ref_sm).abs() / ref_sm).max()) < max(20 * e_tc * scale, 1e-4)

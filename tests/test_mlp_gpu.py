"""Tensor-core MLP (tcgen05 GEMM + InstanceNorm + LeakyReLU + softmax) against the fp32 PyTorch
reference of the same op.  Tolerance: bf16 activations / weights with fp32 accumulation -- logits to
5e-2 absolute on O(1) values, softmax weights to 5 % relative where they matter."""
import numpy as np
import pytest
import torch

from fepe_b200 import _lib, synth
from fepe_b200.models import DeepFNet, ErrorEstimator

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N,K,Co", [(1, 128, 64, 64), (2, 100, 64, 128), (3, 1000, 128, 1024), (2, 1000, 1024, 512),
                                      (2, 333, 512, 256)])
def test_gemm_against_torch_fp32(B, N, K, Co):
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

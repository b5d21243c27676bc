"""Bring-up check of the tcgen05 GEMM (run under a short timeout on the GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import _lib
lib = _lib.lib()
torch.manual_seed(0)
for (B, N, K, Co) in [(1, 128, 64, 64), (2, 100, 64, 128), (3, 1000, 128, 1024), (2, 1000, 1024, 512), (2, 333, 512, 256)]:
    Npad = (N + 127) // 128 * 128
    X = torch.randn(B, Npad, K, device="cuda").bfloat16()
    W = (torch.randn(Co, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(Co, device="cuda")
    Y = torch.full((B * Npad, Co), 7.0, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(B, Co, 2, device="cuda")
    st = lib.fepe_mlp_gemm(X.data_ptr(), W.data_ptr(), bias.data_ptr(), Y.data_ptr(), stats.data_ptr(), B, Npad, N, K, Co,
                           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = (X.float().reshape(-1, K) @ W.float().t() + bias).reshape(B, Npad, Co)
    ref[:, N:] = 0
    Yf = Y.float().reshape(B, Npad, Co)
    err = (Yf - ref).abs().max().item()
    refb = ref.bfloat16().float()
    s1 = refb[:, :N].sum(1); s2 = (refb[:, :N] ** 2).sum(1)
    e1 = (stats[..., 0] - s1).abs().max().item() / (s1.abs().max().item() + 1e-6)
    e2 = (stats[..., 1] - s2).abs().max().item() / (s2.abs().max().item() + 1e-6)
    print(f"B={B} N={N} K={K} Co={Co}: status {st} max|Y-ref| {err:.4f} (ref max {ref.abs().max().item():.2f}) stats rel err {e1:.2e} {e2:.2e}", flush=True)

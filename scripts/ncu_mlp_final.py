"""ncu target: the shipped configuration of the three fat ErrorEstimator layers at B=512, one launch each:
128->1024 (plain persistent GEMM), 1024->512 and 512->256 (fused norm)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import _lib
lib = _lib.lib()
B, N, Npad = 512, 1000, 1024
s_ = torch.cuda.current_stream().cuda_stream
for K, Co, fused in [(128, 1024, False), (1024, 512, True), (512, 256, True)]:
    Yp = torch.randn(B * Npad, K, device="cuda").bfloat16(); W = (torch.randn(Co, K, device="cuda") / K ** 0.5).bfloat16()
    Y = torch.empty(B * Npad, Co, device="cuda", dtype=torch.bfloat16)
    ps = torch.rand(B, K, 2, device="cuda") * 1000 + 1000; g = torch.ones(K, device="cuda"); be = torch.zeros(K, device="cuda")
    ss = torch.empty(B, K // 2, 4, device="cuda"); stats = torch.zeros(B, Co, 2, device="cuda")
    if fused:
        lib.fepe_mlp_scale_shift(ps.data_ptr(), g.data_ptr(), be.data_ptr(), ss.data_ptr(), B, K, N, 1e-5, 0, s_)
        assert lib.fepe_mlp_gemm_norm(Yp.data_ptr(), ss.data_ptr(), 0.01, W.data_ptr(), 0, Y.data_ptr(), stats.data_ptr(), B, Npad, N, K, Co, s_) == 0
    else:
        assert lib.fepe_mlp_gemm(Yp.data_ptr(), W.data_ptr(), 0, Y.data_ptr(), stats.data_ptr(), B, Npad, N, K, Co, s_) == 0
    torch.cuda.synchronize()

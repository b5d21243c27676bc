"""ncu target: one launch each of the persistent MLP GEMM on the ErrorEstimator's layer-2 and layer-3 shapes (B=512)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import _lib
lib = _lib.lib()
B, N, Npad = 512, 1000, 1024
for K, Co in [(128, 1024), (1024, 512)]:
    X = torch.randn(B * Npad, K, device="cuda").bfloat16(); W = (torch.randn(Co, K, device="cuda") / K ** 0.5).bfloat16()
    Y = torch.empty(B * Npad, Co, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(B, Co, 2, device="cuda")
    for _ in range(2):
        assert lib.fepe_mlp_gemm(X.data_ptr(), W.data_ptr(), 0, Y.data_ptr(), stats.data_ptr(), B, Npad, N, K, Co, torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()

"""ncu target: the two tensor-core kernels added late in round 2 -- the matcher's distance tiles (16 pairs x 1200 x 1200 x
256) and the weight-gradient GEMM with 128 x 256 tiles (dW 512 x 1024, M = 16 384 rows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, _lib

g = torch.Generator(device="cuda").manual_seed(0)
B, N = 16, 1200
d1 = torch.nn.functional.normalize(torch.randn(B, N, 256, device="cuda", generator=g), dim=2)
d2 = torch.nn.functional.normalize(d1[:, torch.randperm(N, device="cuda")] + 0.02 * torch.randn(B, N, 256, device="cuda", generator=g), dim=2)
for _ in range(2):
    ops.nn_match_two_way(d1, d2, 1.0)
lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
M, Co, Ci, Npad = 16384, 512, 1024, 1024
dY = torch.randn(M, Co, device="cuda") * 1e-3
Yp = torch.randn(M, Ci, device="cuda")
ss = torch.stack((torch.rand(16, Ci, device="cuda") + 0.5, torch.randn(16, Ci, device="cuda")), 2).contiguous()
amax = dY.abs().max().reshape(1).view(torch.int32).clone()
dW = torch.zeros(Co, Ci, device="cuda")
for _ in range(2):
    lib.fepe_mlp32_wgrad(dY.data_ptr(), amax.data_ptr(), Yp.data_ptr(), ss.data_ptr(), 0.01, dW.data_ptr(), M, Npad, Co, Ci, st)
torch.cuda.synchronize()

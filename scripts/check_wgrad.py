"""Bring-up check of the MN-major tcgen05 weight-gradient GEMM."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import _lib
lib = _lib.lib()
torch.manual_seed(0)
for (M, Co, Ci) in [(64, 128, 64), (128, 128, 128), (1024, 1024, 128), (65536, 512, 1024), (4096, 256, 512)]:
    dY = (torch.randn(M, Co, device="cuda") / 8).bfloat16()
    X = torch.randn(M, Ci, device="cuda").bfloat16()
    dW = torch.zeros(Co, Ci, device="cuda")
    st = lib.fepe_mlp_wgrad(dY.data_ptr(), X.data_ptr(), dW.data_ptr(), M, Co, Ci, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = dY.float().t() @ X.float()
    err = (dW - ref).abs().max().item()
    print(f"M={M} Co={Co} Ci={Ci}: status {st} max|dW-ref| {err:.4e} (ref max {ref.abs().max().item():.2f}, rel {err/ref.abs().max().item():.2e})", flush=True)

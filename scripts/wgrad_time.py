"""Timing of the weight-gradient GEMM per layer shape and tile width (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import _lib

lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
for B in (16, 64, 512):
    Npad = 1024
    M = B * Npad
    for Co, Ci in ((256, 512), (512, 1024), (1024, 128), (128, 64)):
        dY = torch.randn(M, Co, device="cuda") * 1e-3
        Yp = torch.randn(M, Ci, device="cuda")
        ss = torch.stack((torch.rand(B, Ci, device="cuda") + 0.5, torch.randn(B, Ci, device="cuda")), 2).contiguous()
        amax = dY.abs().max().reshape(1).view(torch.int32).clone()
        dW = torch.zeros(Co, Ci, device="cuda")
        res = {}
        for tile in ("128", "256"):
            if tile == "256" and Ci % 256:
                continue
            _lib.set_dispatch("wgrad", tile)
            for _ in range(3):
                lib.fepe_mlp32_wgrad(dY.data_ptr(), amax.data_ptr(), Yp.data_ptr(), ss.data_ptr(), 0.01, dW.data_ptr(), M, Npad, Co, Ci, st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                lib.fepe_mlp32_wgrad(dY.data_ptr(), amax.data_ptr(), Yp.data_ptr(), ss.data_ptr(), 0.01, dW.data_ptr(), M, Npad, Co, Ci, st)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            res[tile] = us
            print(f"B={B:4d} dW {Co:4d}x{Ci:4d} tile {tile}: {us:8.1f} us  {2*M*Co*Ci/us/1e6:7.1f} TFLOP/s algorithmic ({6*M*Co*Ci/us/1e6:7.1f} executed)", flush=True)
        _lib.set_dispatch("wgrad", "auto")

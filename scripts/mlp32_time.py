"""Per-launch device times of one ErrorEstimator evaluation on the fp32-parity tensor-core path (CUDA events around every
C-ABI call), B pairs x N correspondences.  Usage: python scripts/mlp32_time.py [B] [N] [cin]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch

from fepe_b200 import _lib, mlp32
from fepe_b200.models import ErrorEstimator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cin = int(sys.argv[3]) if len(sys.argv) > 3 else 7
torch.manual_seed(0)
ee = ErrorEstimator(cin).cuda()
x = torch.rand(B, N, cin, device="cuda")
lib = _lib.lib()
names, events = [], []
real = {}
for fn in ("fepe_mlp32_first", "fepe_mlp32_scale_shift", "fepe_mlp32_gemm", "fepe_mlp32_last"):
    real[fn] = getattr(lib, fn)


class Timed:
    def __init__(self, name):
        self.name = name

    def __call__(self, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = real[self.name](*a)
        e1.record()
        tag = self.name
        if self.name == "fepe_mlp32_gemm":
            tag += f" K={a[13]} Co={a[14]}"
        names.append(tag)
        events.append((e0, e1))
        return st


class Proxy:
    def __getattr__(self, k):
        return Timed(k) if k in real else getattr(lib, k)


with torch.no_grad():
    for _ in range(3):
        ee.forward_parts(None, None, [x], B, N)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ee.forward_parts(None, None, [x], B, N)
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / 5
    orig = _lib.lib
    _lib.lib = lambda: Proxy()
    mlp32._lib.lib = _lib.lib
    ee.forward_parts(None, None, [x], B, N)
    torch.cuda.synchronize()
    _lib.lib = orig
flop = 2 * B * N * (cin * 64 + 64 * 128 + 128 * 1024 + 1024 * 512 + 512 * 256 + 256)
print(f"ErrorEstimator({cin}) B={B} N={N}: {total:.3f} ms per evaluation = {flop / total / 1e9:.1f} TFLOP/s algorithmic "
      f"({3 * flop / total / 1e9:.1f} TFLOP/s of fp16 MMAs executed)")
for n, (a, b) in zip(names, events):
    ms = a.elapsed_time(b)
    extra = ""
    if "gemm" in n:
        K, Co = int(n.split("K=")[1].split()[0]), int(n.split("Co=")[1])
        f = 2.0 * B * ((N + 127) // 128 * 128) * K * Co
        extra = f"  {f / ms / 1e9:7.1f} TFLOP/s algorithmic, {3 * f / ms / 1e9:7.1f} executed; out {B * 1024 * Co * 4 / 1e6:.0f} MB"
    print(f"  {n:34s} {ms * 1e3:9.1f} us{extra}")

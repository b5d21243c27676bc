"""Do the Gram kernel (12 warps) and the solve + residual kernels really share the SMs?  Times each alone and both
launched on two streams (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth, _lib

B, N = 13320, 1000
base = synth.make_batch(512, N, seed=1, weight_mode="softmax")
aff = ops.hw_affine(base["image_size"])
def mk():
    m = torch.from_numpy(base["matches_xy_ori"]).cuda().repeat((B + 511) // 512, 1, 1)[:B].contiguous()
    w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N).repeat((B + 511) // 512, 1)[:B].contiguous()
    return m, w
mA, wA = mk(); mB, wB = mk()
_lib.set_dispatch("fit", "split"); _lib.set_dispatch("split_rounds", "15")
_lib.set_dispatch("split_pipe", "overlap")
outA = ops.fit_forward(mA, wA, aff); outB = ops.fit_forward(mB, wB, aff)   # valid state in both output sets
torch.cuda.synchronize()
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()

def go(a, b, iters=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    sA.wait_stream(torch.cuda.current_stream()); sB.wait_stream(torch.cuda.current_stream())
    for _ in range(iters):
        if a:
            _lib.set_dispatch("split_pipe", a)
            with torch.cuda.stream(sA):
                ops.fit_forward(mA, wA, aff, out=outA)
        if b:
            _lib.set_dispatch("split_pipe", b)
            with torch.cuda.stream(sB):
                ops.fit_forward(mB, wB, aff, out=outB)
    torch.cuda.current_stream().wait_stream(sA); torch.cuda.current_stream().wait_stream(sB)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for name, a, b in [("K1(12 warps) alone", "d10", None), ("K1(16 warps) alone", "d7", None), ("K2+K3 alone", None, "d8"),
                   ("K1(12) || K2+K3", "d10", "d8"), ("K1(16) || K2+K3", "d7", "d8"), ("K1(12) then K2+K3 same stream", "overlap", None)]:
    go(a, b, 3)
    print(f"{name:32s} {go(a, b):8.1f} us per iteration (B={B})", flush=True)

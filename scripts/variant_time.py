"""Time forward-kernel build variants (development aid).  usage: variant_time.py lib.so [kernel]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
os.environ["FEPE_B200_LIB"] = os.path.abspath(sys.argv[1])
if len(sys.argv) > 2 and sys.argv[2] in ("ring", "small"):
    os.environ["FEPE_FIT_KERNEL"] = sys.argv[2]
import torch
from fepe_b200 import ops, synth

def timeit(B, N, iters=20):
    base = synth.make_batch(256, N, seed=1, weight_mode="softmax")
    m = torch.from_numpy(base["matches_xy_ori"]).cuda().repeat((B + 255) // 256, 1, 1)[:B].contiguous()
    w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N).repeat((B + 255) // 256, 1)[:B].contiguous()
    aff = ops.hw_affine(base["image_size"])
    out = (torch.empty(B, 3, 3, device="cuda"), torch.empty(B, N, device="cuda"), torch.empty(B, N, device="cuda"), None)
    for _ in range(3):
        ops.fit_forward(m, w, aff, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.fit_forward(m, w, aff, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms * 1e3, B / ms * 1e3

def time_bwd(B, N, iters=10):
    base = synth.make_batch(256, N, seed=1, weight_mode="softmax")
    m = torch.from_numpy(base["matches_xy_ori"]).cuda().repeat((B + 255) // 256, 1, 1)[:B].contiguous()
    w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N).repeat((B + 255) // 256, 1)[:B].contiguous()
    aff = ops.hw_affine(base["image_size"])
    F, res, epi, saved = ops.fit_forward(m, w, aff, want_saved=True)
    gF, gr, ge = torch.randn_like(F), torch.randn_like(res), torch.randn_like(epi)
    for _ in range(2):
        ops.fit_backward(m, w, saved, gF, gr, ge, aff)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.fit_backward(m, w, saved, gF, gr, ge, aff)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms * 1e3, B / ms * 1e3

res = []
for B, N in [(256, 1000), (32768, 1000), (16384, 2000)]:
    us, pps = timeit(B, N)
    res.append(f"B={B} N={N}: {us:8.1f} us {pps/1e6:7.2f} Mpairs/s")
if "--bwd" in sys.argv:
    for B, N in [(256, 1000), (32768, 1000)]:
        us, pps = time_bwd(B, N)
        res.append(f"BWD B={B}: {us:8.1f} us {pps/1e6:7.2f} Mpairs/s")
print(os.path.basename(sys.argv[1]), sys.argv[2:] , " | ".join(res), flush=True)

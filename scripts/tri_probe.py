"""Development aid: where the time goes after the tridiagonal solver (small-kernel phases, per-kernel times of the split
pipeline via the torch profiler's CUDA activity, ring kernel at a mid-size batch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200"), os.path.join(ROOT, "scripts")]
import torch
import small_phases  # noqa: F401  (prints the latency kernel's phases)
from split_time import run
from torch.profiler import profile, ProfilerActivity
from fepe_b200 import ops, synth

run(2048, 1000, "ring", iters=20)
run(32768, 1000, "split", iters=20)
B, N = 32768, 1000
base = synth.make_batch(512, N, seed=1, weight_mode="softmax")
m = torch.from_numpy(base["matches_xy_ori"]).cuda().repeat(B // 512, 1, 1).contiguous()
w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N).repeat(B // 512, 1).contiguous()
aff = ops.hw_affine(base["image_size"])
out = ops.fit_forward(m, w, aff)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ops.fit_forward(m, w, aff, out=out)
    torch.cuda.synchronize()
for e in prof.key_averages():
    print(f"{e.key[:60]:60s} n={e.count:3d} mean {e.device_time_total / max(e.count, 1):9.1f} us")

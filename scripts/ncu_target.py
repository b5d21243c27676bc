"""Target for ncu captures: a few launches of the fused forward kernel at the config batch and at a
saturating batch (development aid; numbers printed under ncu are never bench values)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
base = synth.make_batch(256, N, seed=1, weight_mode="softmax")
aff = ops.hw_affine(base["image_size"])
m = torch.from_numpy(base["matches_xy_ori"]).cuda()
w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N)
for B in (256, 32768):
    mm = m.repeat(B // 256, 1, 1).contiguous()
    ww = w.repeat(B // 256, 1).contiguous()
    for _ in range(2):
        ops.fit_forward(mm, ww, aff)
    torch.cuda.synchronize()

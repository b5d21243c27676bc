"""ncu target: saturating batch of the forward kernel (lib / kernel chosen by env)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth
N, B = 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 32768
base = synth.make_batch(256, N, seed=1, weight_mode="softmax")
aff = ops.hw_affine(base["image_size"])
m = torch.from_numpy(base["matches_xy_ori"]).cuda().repeat(B // 256, 1, 1).contiguous()
w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N).repeat(B // 256, 1).contiguous()
for _ in range(3):
    ops.fit_forward(m, w, aff)
torch.cuda.synchronize()

"""Target for ncu captures of the split forward pipeline at a saturating batch (inference: fp32 Gram + refinement)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
saved = len(sys.argv) > 3 and sys.argv[3] == "saved"
base = synth.make_batch(512, N, seed=1, weight_mode="softmax")
aff = ops.hw_affine(base["image_size"])
m = torch.from_numpy(base["matches_xy_ori"]).cuda().repeat(B // 512, 1, 1).contiguous()
w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N).repeat(B // 512, 1).contiguous()
for _ in range(3):
    out = ops.fit_forward(m, w, aff, want_saved=saved)
torch.cuda.synchronize()

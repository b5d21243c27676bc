"""Target for ncu captures of the split forward pipeline (development aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
os.environ["FEPE_FIT_KERNEL"] = "split"
import torch
from fepe_b200 import ops, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
base = synth.make_batch(512, N, seed=1, weight_mode="softmax")
aff = ops.hw_affine(base["image_size"])
m = torch.from_numpy(base["matches_xy_ori"]).cuda().repeat(B // 512, 1, 1).contiguous()
w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N).repeat(B // 512, 1).contiguous()
for _ in range(2):
    out = ops.fit_forward(m, w, aff, want_saved=True)
torch.cuda.synchronize()
its = out[3][:, 52]
print("iterations per pair: mean %.2f max %d; mean over warps of the per-warp max %.2f" % (
    float(its.mean()), int(its.max()), float(its.reshape(-1, 32).max(dim=1).values.mean())))
print("histogram:", torch.bincount(its.long()).tolist())
ph = out[3][:, 56:60].mean(0).cpu().numpy()
print("K1 mean cycles per pair-team: wait %.0f hartley %.0f gram %.0f reduce+store %.0f" % tuple(ph))

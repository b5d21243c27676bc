"""Per-phase cycles of the latency kernel at the config batch (development aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth
B, N = 256, 1000
base = synth.make_batch(B, N, seed=1, weight_mode="softmax")
m = torch.from_numpy(base["matches_xy_ori"]).cuda(); w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N)
aff = ops.hw_affine(base["image_size"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    out = ops.fit_forward(m, w, aff, want_saved=True)
torch.cuda.synchronize()
ph = out[3][:, 56:63]
names = ["copy", "hartley", "gram+reduce", "solve", "residual", "-", "eig only"]
print("small kernel, cold inputs: mean / max cycles per pair")
for i, n in enumerate(names):
    print(f"  {n:12s} {float(ph[:, i].mean()):9.0f} {float(ph[:, i].max()):9.0f}")
print("  total mean", float(ph[:, :5].sum(1).mean()), "max", float(ph[:, :5].sum(1).max()), "rounds mean", float(out[3][:, 52].mean()))

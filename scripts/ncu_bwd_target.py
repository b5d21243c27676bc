"""ncu target: forward (saving state) + backward of the fit at a config batch and at a saturating batch, and one
ErrorEstimator training step (forward + backward kernels of the fp32-parity MLP)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch

from fepe_b200 import ops, synth
from fepe_b200.models import ErrorEstimator

for B in (256, 8192):
    d = synth.make_batch(min(B, 512), 1000, seed=3, weight_mode="softmax")
    m = torch.from_numpy(d["matches_xy_ori"]).cuda().repeat(max(1, B // 512), 1, 1).contiguous()
    w = torch.from_numpy(d["weights"]).cuda().reshape(-1, 1000).repeat(max(1, B // 512), 1).contiguous()
    aff = ops.hw_affine(d["image_size"])
    gF = torch.randn(m.shape[0], 3, 3, device="cuda")
    gr = torch.randn(m.shape[0], 1000, device="cuda") * 1e-3
    ge = torch.randn(m.shape[0], 1000, device="cuda") * 1e-3
    for _ in range(2):
        F, res, epi, saved = ops.fit_forward(m, w, aff, want_saved=True)
        ops.fit_backward(m, w, saved, gF, gr, ge, aff, 0.5)
torch.manual_seed(0)
ee = ErrorEstimator(7).cuda()
x = torch.rand(64, 7, 1000, device="cuda", requires_grad=True)
for _ in range(2):
    ee.zero_grad()
    (ee(x) * torch.randn(64, 1, 1000, device="cuda")).sum().backward()
torch.cuda.synchronize()

"""ncu target: one ErrorEstimator training evaluation (forward + backward) at the batch of a training step (16 pairs x
1000 correspondences) -- the kernels that were re-sized for small batches (cluster-split last layer, InstanceNorm /
last-layer backward row tiles)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch

from fepe_b200.models import ErrorEstimator

torch.manual_seed(0)
ee = ErrorEstimator(7).cuda()
x = torch.rand(16, 7, 1000, device="cuda", requires_grad=True)
g = torch.randn(16, 1, 1000, device="cuda")
for _ in range(3):
    ee.zero_grad()
    (ee(x) * g).sum().backward()
torch.cuda.synchronize()

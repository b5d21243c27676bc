"""Development aid: one tensor-core ErrorEstimator evaluation at B=512 (run under `ncu --metrics gpu__time_duration.sum`)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200.models import ErrorEstimator
B, N = 512, 1000
ee = ErrorEstimator(4).cuda()
ee.tensor_cores = True
x = torch.rand(B, 4, N, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = ee(x)
torch.cuda.synchronize()

"""ncu target: three ErrorEstimator(7) evaluations at B x N on the fp32-parity tensor-core path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch

from fepe_b200.models import ErrorEstimator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
ee = ErrorEstimator(7).cuda()
x = torch.rand(B, 1000, 7, device="cuda")
with torch.no_grad():
    for _ in range(3):
        ee.forward_parts(None, None, [x], B, 1000)
torch.cuda.synchronize()

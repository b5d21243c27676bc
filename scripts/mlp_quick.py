"""ErrorEstimator / DeepFNet forward timing only (B=512, N=1000), tensor-core path with the shipped defaults."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import synth
from fepe_b200.models import DeepFNet, ErrorEstimator

def ev(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

N = 1000
FLOP_PER_PT = 2 * (4 * 64 + 64 * 128 + 128 * 1024 + 1024 * 512 + 512 * 256 + 256)
for cin in (4, 7):
    ee = ErrorEstimator(cin).cuda(); ee.tensor_cores = True
    x = torch.rand(512, cin, N, device="cuda")
    with torch.no_grad():
        t = ev(lambda: ee(x))
    print(f"ErrorEstimator({cin}) B=512: {t:.3f} ms  {512*N*FLOP_PER_PT/t/1e9:.1f} TFLOP/s", flush=True)
net = DeepFNet(depth=5, image_size=[376, 1241, 3], if_quality=False).cuda()
net.enable_tensor_core_mlp()
d = synth.make_batch(64, N, seed=1)
m = torch.from_numpy(d["matches_xy_ori"]).cuda().repeat(8, 1, 1).contiguous()
with torch.no_grad():
    t = ev(lambda: net({"matches_xy_ori": m}), iters=5, warm=2)
print(f"DeepFNet forward depth 5 B=512: {t:.2f} ms  {512/t*1e3:.0f} pairs/s", flush=True)

"""compute-sanitizer target: the split pipeline in chunked / overlapped launch modes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth, _lib
B, N = int(sys.argv[1]), 1000
pipe, rounds, iters = sys.argv[2], sys.argv[3], int(sys.argv[4])
base = synth.make_batch(512, N, seed=1, weight_mode="softmax")
m = torch.from_numpy(base["matches_xy_ori"]).cuda().repeat((B + 511) // 512, 1, 1)[:B].contiguous()
w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N).repeat((B + 511) // 512, 1)[:B].contiguous()
aff = ops.hw_affine(base["image_size"])
_lib.set_dispatch("fit", "split"); _lib.set_dispatch("split_pipe", pipe); _lib.set_dispatch("split_rounds", rounds)
out = ops.fit_forward(m, w, aff)
for i in range(iters):
    ops.fit_forward(m, w, aff, out=out)
    torch.cuda.synchronize()
print("ok", pipe, rounds, float(out[0].abs().sum()))

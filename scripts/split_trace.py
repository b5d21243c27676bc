"""Timeline of the split pipeline's kernels from the per-CTA trace (fepe_debug_trace): when each launch ran, and how many
CTAs of the solve / residual kernels ran on an SM while a Gram CTA was resident there (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import numpy as np
import torch
from fepe_b200 import ops, synth, _lib

B, N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768, 1000
pipe, rounds = "whole", "-"
base = synth.make_batch(512, N, seed=1, weight_mode="softmax")
m = torch.from_numpy(base["matches_xy_ori"]).cuda().repeat((B + 511) // 512, 1, 1)[:B].contiguous()
w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N).repeat((B + 511) // 512, 1)[:B].contiguous()
aff = ops.hw_affine(base["image_size"])
_lib.set_dispatch("fit", "split")
out = ops.fit_forward(m, w, aff)
for _ in range(2):
    ops.fit_forward(m, w, aff, out=out)
torch.cuda.synchronize()
cap = 1 << 17
buf = torch.zeros(cap, 4, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib().fepe_debug_trace(buf.data_ptr(), cap), "trace")
ops.fit_forward(m, w, aff, out=out)
torch.cuda.synchronize()
n = min(_lib.lib().fepe_debug_trace_count(), cap)
_lib.check(_lib.lib().fepe_debug_trace(None, 0), "trace off")
r = buf[:n].cpu().numpy()
t0 = r[:, 2].min()
r[:, 2] -= t0; r[:, 3] -= t0
print(f"{pipe} rounds={rounds} B={B}: {n} CTA records, span {r[:,3].max()/1e3:.1f} us")
for kid, name in ((1, "gram"), (2, "solve"), (3, "resid")):
    k = r[r[:, 0] == kid]
    if len(k) == 0:
        continue
    # split into launches by gaps in start time
    order = np.argsort(k[:, 2]); k = k[order]
    print(f"  {name}: {len(k)} CTAs, mean CTA time {np.mean(k[:,3]-k[:,2])/1e3:.1f} us, first start {k[0,2]/1e3:.1f}, last end {k[:,3].max()/1e3:.1f} us")
g = r[r[:, 0] == 1]
others = r[r[:, 0] != 1]
# co-residency: for every solve / resid CTA, was a Gram CTA resident on the same SM for its whole life?
co = 0
by_sm = {}
for row in g:
    by_sm.setdefault(int(row[1]), []).append((row[2], row[3]))
for row in others:
    for (a, b) in by_sm.get(int(row[1]), []):
        if a <= row[2] and row[3] <= b:
            co += 1
            break
print(f"  solve/resid CTAs that ran entirely while a Gram CTA was resident on their SM: {co} of {len(others)}")
# coarse timeline: 20 us bins, number of CTAs of each kernel active
T = r[:, 3].max()
bins = np.arange(0, T + 20000, 20000)
for kid, name in ((1, "gram"), (2, "solve"), (3, "resid")):
    k = r[r[:, 0] == kid]
    act = [int(np.sum((k[:, 2] < b + 20000) & (k[:, 3] > b))) for b in bins[:-1]]
    print(f"  active {name:5s} CTAs per 20 us bin: {act}")

"""Timing of the split forward pipeline against the fused ring kernel (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth
from fepe_b200 import _lib as _fepe_lib


def run(B, N, kern, iters=10, saved=False, env=None):
    _fepe_lib.set_dispatch("fit", kern)
    for k, v in (env or {}).items():
        _fepe_lib.set_dispatch(k, v)
    base = synth.make_batch(512, N, seed=1, weight_mode="softmax")
    m = torch.from_numpy(base["matches_xy_ori"]).cuda()
    w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N)
    reps = (B + 511) // 512
    m = m.repeat(reps, 1, 1)[:B].contiguous(); w = w.repeat(reps, 1)[:B].contiguous()
    aff = ops.hw_affine(base["image_size"])
    out = None
    for _ in range(3):
        out = ops.fit_forward(m, w, aff, want_saved=saved)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.fit_forward(m, w, aff, want_saved=saved, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    by = B * (N * 28 + 36)
    print(f"{kern:6s} {env or ''} B={B:6d} N={N:5d} saved={saved}: {ms*1e3:9.1f} us  {B/ms*1e3/1e6:7.2f} M pairs/s  "
          f"{by/ms/1e6:7.1f} GB/s ({by/ms/1e6/6574.5*100:4.1f}% of HBM peak)", flush=True)
    for k in (env or {}):
        _fepe_lib.set_dispatch(k, "auto")
    return out


if __name__ == "__main__":
    for B, N in [(32768, 1000), (16384, 2000), (4096, 1000), (1024, 1000), (256, 1000)]:
        a = run(B, N, "ring")
        for t in ("1", "2", "4"):
            if N == 1000 and t == "1":
                continue
            b = run(B, N, "split", env={"gram_team": t})
        relF = ((a[0] - b[0]).flatten(1).norm(dim=1) / a[0].flatten(1).norm(dim=1)).max().item()
        print(f"   max rel F difference split vs ring: {relF:.2e}; resid {float((a[1]-b[1]).abs().max()):.2e}")
    run(32768, 1000, "split", saved=True)
    run(32768, 1000, "ring", saved=True)

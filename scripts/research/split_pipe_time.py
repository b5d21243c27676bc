"""Timing of the split pipeline's launch modes (whole batch / L2-sized chunks / chunks with the solve + residual kernels
overlapped under the next Gram kernel) at the saturating batch (development aid; `python scripts/split_pipe_time.py`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth
from fepe_b200 import _lib


def run(B, N, pipe, rounds="auto", team="auto", iters=10, saved=False, ref=None):
    _lib.set_dispatch("fit", "split"); _lib.set_dispatch("split_pipe", pipe); _lib.set_dispatch("split_rounds", rounds)
    _lib.set_dispatch("gram_team", team)
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
    diff = ""
    if ref is not None:
        relF = ((ref[0] - out[0]).flatten(1).norm(dim=1) / ref[0].flatten(1).norm(dim=1)).max().item()
        diff = f"  vs whole: F {relF:.1e} resid {float((ref[1]-out[1]).abs().max()):.1e} epi {float((ref[2]-out[2]).abs().max()):.1e}"
    print(f"{pipe:8s} rounds={rounds:4s} team={team:4s} B={B:6d} N={N:5d} saved={saved}: {ms*1e3:9.1f} us  {B/ms*1e3/1e6:7.2f} M pairs/s  "
          f"{by/ms/1e6:7.1f} GB/s ({by/ms/1e6/6574.5*100:4.1f}% of HBM peak){diff}", flush=True)
    for k in ("fit", "split_pipe", "split_rounds", "gram_team"):
        _lib.set_dispatch(k, "auto")
    return [o.clone() if o is not None else None for o in out]


if __name__ == "__main__":
    for B, N in [(32768, 1000)]:
        ref = run(B, N, "whole")
        for r in ("2", "4", "6", "8", "12"):
            run(B, N, "chunks", r, ref=ref)
        for r in ("1", "2", "3", "4", "6", "8", "12"):
            run(B, N, "overlap", r, ref=ref)
        ref = run(B, N, "whole")

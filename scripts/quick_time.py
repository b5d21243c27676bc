"""Ad-hoc timing of the forward kernel (development aid; bench.py is the contract)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth, _lib

def timeit(B, N, iters=20, epi=True, saved=False):
    base = synth.make_batch(min(B, 512), N, seed=1, weight_mode="softmax")
    m = torch.from_numpy(base["matches_xy_ori"]).cuda()
    w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N)
    reps = (B + m.shape[0] - 1) // m.shape[0]
    m = m.repeat(reps, 1, 1)[:B].contiguous(); w = w.repeat(reps, 1)[:B].contiguous()
    aff = ops.hw_affine(base["image_size"])
    for _ in range(3):
        ops.fit_forward(m, w, aff, want_epi=epi, want_saved=saved)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.fit_forward(m, w, aff, want_epi=epi, want_saved=saved)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bytes_ = B * (N * (28 if epi else 24) + 36)
    print(f"B={B:6d} N={N:5d} epi={epi} saved={saved}: {ms*1e3:9.1f} us/launch  {B/ms*1e3:12.0f} pairs/s  "
          f"{bytes_/ms/1e6:8.1f} GB/s", flush=True)

if __name__ == "__main__":
    for kern in ("ring", "small"):
        _lib.set_dispatch("fit", kern)
        print("kernel", kern)
        timeit(256, 1000)
        timeit(64, 2000)
        timeit(148, 1000)
    _lib.set_dispatch("fit", "auto")
    for B, N in [(256, 1000), (32768, 1000), (64, 2000), (16384, 2000)]:
        timeit(B, N)

    # per-phase cycles from the saved diagnostics
    for B, N in [(256, 1000), (32768, 1000)]:
        base = synth.make_batch(256, N, seed=1, weight_mode="softmax")
        m = torch.from_numpy(base["matches_xy_ori"]).cuda().repeat(B // 256, 1, 1).contiguous()
        w = torch.from_numpy(base["weights"]).cuda().reshape(-1, N).repeat(B // 256, 1).contiguous()
        _, _, _, sv = ops.fit_forward(m, w, ops.hw_affine(base["image_size"]), want_saved=True)
        torch.cuda.synchronize()
        ph = sv[:, 56:63].mean(0).cpu().numpy()
        print(f"B={B} N={N} mean cycles/pair: wait {ph[0]:.0f} hartley {ph[1]:.0f} gram {ph[2]:.0f} solve {ph[3]:.0f} "
              f"(reduce {ph[5]:.0f} eig {ph[6]:.0f}) resid {ph[4]:.0f}; factorisations mean {float(sv[:,52].mean()):.2f}")

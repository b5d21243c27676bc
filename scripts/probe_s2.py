"""Session-2 probes (development aid): residual-kernel traversal order, H2D copy bandwidth at the step size,
latency-kernel phases."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200"), os.path.join(ROOT, "scripts")]
import torch
from split_time import run

for rev in ("0", "1"):
    for B in (32768, 8192, 4096):
        run(B, 1000, "split", iters=20, env={"FEPE_RESID_REVERSE": rev})
run(16384, 2000, "split", iters=10, env={"FEPE_RESID_REVERSE": "0"})
run(16384, 2000, "split", iters=10, env={"FEPE_RESID_REVERSE": "1"})

# ---- H2D bandwidth at the size of one C2 step (5.77 MB) and at 256 MB
for nbytes in (5767168, 256 << 20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
    for nstreams in (1, 2):
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        reps = 200 if nbytes < (64 << 20) else 10
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(reps):
            with torch.cuda.stream(streams[i % nstreams]):
                d[i % 2].copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"H2D {nbytes/1e6:8.2f} MB x {reps} on {nstreams} stream(s): {nbytes*reps/dt/1e9:6.1f} GB/s, {dt/reps*1e6:8.1f} us per copy", flush=True)
    # split one copy into 2 halves on 2 streams
    streams = [torch.cuda.Stream() for _ in range(2)]
    half = nbytes // 2
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                d[i % 2][k * half:(k + 1) * half].copy_(h[k * half:(k + 1) * half], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D {nbytes/1e6:8.2f} MB x {reps} as 2 halves on 2 streams: {nbytes*reps/dt/1e9:6.1f} GB/s", flush=True)

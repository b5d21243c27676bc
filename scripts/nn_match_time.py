"""Timing of fepe_nn_match against the numpy algorithm on the host (development aid / profiles)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import numpy as np, torch
from fepe_b200 import ops, _lib
from oracle import nn_match_oracle as NO
for kern, B, N in [(k, b, n) for (b, n) in [(16, 1200), (128, 1000)] for k in ("simt", "tc")]:
    _lib.set_dispatch("nn_dist", kern)
    g = torch.Generator(device="cuda").manual_seed(0)
    d1 = torch.nn.functional.normalize(torch.randn(B, N, 256, device="cuda", generator=g), dim=2)
    d2 = torch.nn.functional.normalize(d1[:, torch.randperm(N, device="cuda")] + 0.02 * torch.randn(B, N, 256, device="cuda", generator=g), dim=2)
    for _ in range(3):
        ops.nn_match_two_way(d1, d2, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = ops.nn_match_two_way(d1, d2, 1.0)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    fl = 2.0 * B * N * N * 256
    a, b = d1[0].cpu().numpy().T.copy(), d2[0].cpu().numpy().T.copy()
    t0 = time.perf_counter()
    for _ in range(3):
        NO.nn_match_two_way(a, b, 1.0)
    cpu = (time.perf_counter() - t0) / 3
    print(f"fepe_nn_match[{kern}] B={B} N1=N2={N} D=256: {us:9.1f} us = {fl/us/1e6:6.1f} TFLOP/s algorithmic, {B/us*1e6:9.0f} pairs/s, "
          f"matches/pair {float(out[3].float().mean()):.0f} | numpy on the host: {cpu*1e3:7.2f} ms per pair", flush=True)

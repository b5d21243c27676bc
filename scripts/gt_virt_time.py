"""Development aid: fepe_gt_virt launch time vs the host path it replaces (cv2.correctMatches per sample)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import numpy as np
import torch
import __graft_entry__ as g
g.build()
from fepe_b200 import gt as G, synth

for B in (16, 128, 1024):
    d = synth.make_batch(B, 8, seed=1)
    Rt, Ks = torch.from_numpy(d["delta_Rtijs_4_4"]).cuda(), torch.from_numpy(d["Ks"]).cuda()
    grids = G.get_virt_x1x2_grid(d["image_size"], device="cuda")
    for _ in range(3):
        out = G.gt_sample_batch(Rt, Ks, d["image_size"], grids=grids)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = G.gt_sample_batch(Rt, Ks, d["image_size"], grids=grids)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    line = f"fepe_gt_virt B={B:5d} x 100 grid points: {us:8.1f} us per batch ({us / B:6.2f} us per sample)"
    try:
        import cv2
        g1, g2 = (t.cpu().numpy() for t in grids)
        n = min(B, 64)
        t0 = time.perf_counter()
        for b in range(n):
            cv2.correctMatches(d["F_gt"][b].astype(np.float64), g2[None], g1[None])
        host = (time.perf_counter() - t0) / n * 1e6
        line += f"; host cv2.correctMatches alone: {host:7.1f} us per sample (1 core)"
    except Exception as e:      # pragma: no cover
        line += f"; cv2 unavailable ({e})"
    print(line)

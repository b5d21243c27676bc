"""Timing of fepe_recover_pose against cv2.recoverPose on the host (development aid / profiles)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import numpy as np, torch
from fepe_b200 import ops, synth
for B, N in [(256, 1000), (64, 2000), (1024, 1000)]:
    d = synth.make_batch(min(B, 256), N, seed=1)
    rep = (B + 255) // 256
    t = lambda k: torch.from_numpy(np.concatenate([d[k]] * rep)[:B]).cuda()
    E, K, m, Rt = t("E_gt"), t("Ks"), t("matches_xy_ori"), t("delta_Rtijs_4_4")
    for _ in range(3):
        ops.recover_pose(E, K, m, Rt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.recover_pose(E, K, m, Rt)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    line = f"fepe_recover_pose B={B} N={N}: {us:8.1f} us per launch = {B/us:6.2f} M pairs/s"
    try:
        import cv2
        n = min(B, 64)
        t0 = time.perf_counter()
        for b in range(n):
            Kb = d["Ks"][b % 256]; mb = d["matches_xy_ori"][b % 256].astype(np.float64)
            cv2.recoverPose(d["E_gt"][b % 256].astype(np.float64), mb[:, :2], mb[:, 2:], focal=float(Kb[0, 0]), pp=(float(Kb[0, 2]), float(Kb[1, 2])))
        cpu = (time.perf_counter() - t0) / n
        line += f" | cv2.recoverPose on one host core: {cpu*1e6:8.1f} us per pair = {1/cpu:8.0f} pairs/s"
    except Exception as ex:
        line += f" | cv2 unavailable ({ex})"
    print(line, flush=True)

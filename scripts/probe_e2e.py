"""Development aid: where does the end-to-end step spend its time?  H2D only / +kernel / +D2H, eager and graph."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth
from fepe_b200.staging import StagedStep

B, N = 256, 1000
dev = torch.device("cuda")
d = synth.make_batch(B, N, seed=1, weight_mode="softmax")
aff = ops.hw_affine(d["image_size"])
V = d["pts1_virt"].shape[1]


def timeit(name, fn, nslots, steps=400):
    for i in range(8):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:58s} slots={nslots}: {dt/steps*1e6:7.1f} us/step  (host issue {t_issue/steps*1e6:6.1f} us/step)  {B*steps/dt/1e6:5.2f} M pairs/s", flush=True)


for nslots in (1, 2, 4):
    stages = [StagedStep(B, N, V, dev) for _ in range(nslots)]
    streams = [torch.cuda.Stream() for _ in range(nslots)]
    for s in stages:
        s.pack(d, out=s.h_in)

    def h2d_only(i):
        j = i % nslots
        with torch.cuda.stream(streams[j]):
            stages[j].d_in.copy_(stages[j].h_in, non_blocking=True)

    def full_eager(i):
        j = i % nslots
        stages[j].run(streams[j], aff)

    timeit("eager: H2D only", h2d_only, nslots)
    timeit("eager: H2D + fit_pose + D2H", full_eager, nslots)
    for s, st in zip(stages, streams):
        s.capture(st, aff)
    timeit("graph: H2D + fit_pose + D2H", lambda i: stages[i % nslots].replay(), nslots)

    # graph without the D2H, and kernel-only graph
    gs = []
    for s, st in zip(stages, streams):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            s.d_in.copy_(s.h_in, non_blocking=True)
        gs.append(g)

    def g_h2d(i):
        j = i % nslots
        with torch.cuda.stream(streams[j]):
            gs[j].replay()
    timeit("graph: H2D only", g_h2d, nslots)

# two-step software pipeline on ONE stream pair: copy stream + compute stream with events (classic double buffer)
nslots = 3
stages = [StagedStep(B, N, V, dev) for _ in range(nslots)]
for s in stages:
    s.pack(d, out=s.h_in)
copy_s, comp_s, out_s = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
ev_in = [torch.cuda.Event() for _ in range(nslots)]
ev_k = [torch.cuda.Event() for _ in range(nslots)]
ev_out = [torch.cuda.Event() for _ in range(nslots)]
from fepe_b200 import _lib


def pipelined(i):
    j = i % nslots
    s = stages[j]
    with torch.cuda.stream(copy_s):
        copy_s.wait_event(ev_k[j])            # the slot's previous kernel has consumed d_in
        s.d_in.copy_(s.h_in, non_blocking=True)
        ev_in[j].record(copy_s)
    with torch.cuda.stream(comp_s):
        comp_s.wait_event(ev_in[j])
        comp_s.wait_event(ev_out[j])          # the slot's previous results have left d_out
        v = lambda k: s._view(s.d_in, k)
        F = s.d_out[:B * 9].view(B, 3, 3)
        pose = s.d_out[B * 9:].view(1, B, _lib.POSE_OUT_FLOATS)
        ops.fit_pose_forward(v("matches_xy_ori"), v("weights"), aff, v("Ks"), v("q_cam"), v("t_cam"),
                             v("delta_Rtijs_4_4"), v("pts1_virt"), v("pts2_virt"), out=(F, s.d_res, s.d_epi, None, pose[0]))
        ev_k[j].record(comp_s)
    with torch.cuda.stream(out_s):
        out_s.wait_event(ev_k[j])
        s.h_out.copy_(s.d_out, non_blocking=True)
        ev_out[j].record(out_s)


timeit("eager: dedicated copy / compute / readback streams", pipelined, nslots)

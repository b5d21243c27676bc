"""Fused norm -> GEMM (fepe_mlp_gemm_norm) against norm kernel + GEMM per ErrorEstimator layer, and the whole
ErrorEstimator / DeepFNet forward with and without the fusion.  `--ncu` runs two fused launches of layer 3 only."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import synth, _lib
from fepe_b200 import _lib as _fepe_lib
import fepe_b200.mlp_tc as mt
from fepe_b200.models import DeepFNet, ErrorEstimator

def ev(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

lib = _lib.lib()
N = 1000; Npad = 1024
FLOP_PER_PT = 2 * (4 * 64 + 64 * 128 + 128 * 1024 + 1024 * 512 + 512 * 256 + 256)
shapes = [(1024, 512)] if "--ncu" in sys.argv else [(64, 128), (128, 1024), (1024, 512), (512, 256)]
for B in ((512,) if "--ncu" in sys.argv else (64, 512)):
    for K, Co in shapes:
        Yp = torch.randn(B * Npad, K, device="cuda").bfloat16(); W = (torch.randn(Co, K, device="cuda") / K ** 0.5).bfloat16()
        X = torch.empty_like(Yp); Y = torch.empty(B * Npad, Co, device="cuda", dtype=torch.bfloat16)
        ps = torch.rand(B, K, 2, device="cuda") * 1000 + 1000; g = torch.ones(K, device="cuda"); be = torch.zeros(K, device="cuda")
        ss = torch.empty(B, K // 2, 4, device="cuda"); stats = torch.zeros(B, Co, 2, device="cuda")
        s_ = torch.cuda.current_stream().cuda_stream
        def unf():
            lib.fepe_mlp_norm(Yp.data_ptr(), ps.data_ptr(), g.data_ptr(), be.data_ptr(), X.data_ptr(), B, Npad, N, K, 1e-5, 0.01, s_)
            lib.fepe_mlp_gemm(X.data_ptr(), W.data_ptr(), 0, Y.data_ptr(), stats.data_ptr(), B, Npad, N, K, Co, s_)
        def fus():
            lib.fepe_mlp_scale_shift(ps.data_ptr(), g.data_ptr(), be.data_ptr(), ss.data_ptr(), B, K, N, 1e-5, 0, s_)
            assert lib.fepe_mlp_gemm_norm(Yp.data_ptr(), ss.data_ptr(), 0.01, W.data_ptr(), 0, Y.data_ptr(), stats.data_ptr(), B, Npad, N, K, Co, s_) == 0
        if "--ncu" in sys.argv:
            fus(); fus(); torch.cuda.synchronize(); sys.exit(0)
        tu = ev(unf)
        _fepe_lib.set_dispatch("mlp_fuse", "1"); tf1 = ev(fus)
        _fepe_lib.set_dispatch("mlp_fuse", "2"); tf = ev(fus)
        print(f"B={B} K={K} Co={Co}: norm + gemm {tu*1e3:7.1f} us | fused v1 {tf1*1e3:7.1f} us | fused v2 {tf*1e3:7.1f} us ({2*B*Npad*K*Co/tf/1e9:7.1f} TFLOP/s)", flush=True)
for fuse, fv in ((False, "2"), (True, "1"), (True, "2")):
    _fepe_lib.set_dispatch("mlp_fuse", fv)
    print(f"-- mlp_fuse dispatch={fv}")
    for B in (64, 512):
        ee = ErrorEstimator(4).cuda(); ee.tensor_cores = True
        x = torch.rand(B, 4, N, device="cuda")
        with torch.no_grad():
            ee(x); ee._tc.fuse_norm = fuse
            t = ev(lambda: ee(x))
        print(f"fuse={fuse} ErrorEstimator B={B}: {t:.3f} ms  {B*N*FLOP_PER_PT/t/1e9:.1f} TFLOP/s", flush=True)
    net = DeepFNet(depth=5, image_size=[376, 1241, 3], if_quality=False).cuda()
    net.enable_tensor_core_mlp()
    d = synth.make_batch(64, N, seed=1)
    m = torch.from_numpy(d["matches_xy_ori"]).cuda().repeat(8, 1, 1).contiguous()
    with torch.no_grad():
        net({"matches_xy_ori": m})
        for mod in net.modules():
            if getattr(mod, "_tc", None) is not None: mod._tc.fuse_norm = fuse
        t = ev(lambda: net({"matches_xy_ori": m}), iters=5, warm=2)
    print(f"fuse={fuse} DeepFNet forward depth 5 B=512: {t:.2f} ms  {512/t*1e3:.0f} pairs/s", flush=True)

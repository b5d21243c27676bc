"""Timing of the weight MLP: cuDNN fp32 (reference arithmetic) vs the tcgen05 bf16 path, and of a
whole DeepFNet forward (config C4 shapes).  Development aid; prints one summary line per case."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import synth, _lib
from fepe_b200.models import DeepFNet, ErrorEstimator

def ev(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

FLOP_PER_PT = 2 * (4 * 64 + 64 * 128 + 128 * 1024 + 1024 * 512 + 512 * 256 + 256)
out = {}
for B, N in [(64, 1000), (512, 1000)]:
    ee = ErrorEstimator(4).cuda()
    x = torch.rand(B, 4, N, device="cuda")
    with torch.no_grad():
        t32 = ev(lambda: ee(x))
        ee.tensor_cores = True
        ttc = ev(lambda: ee(x))
    fl = B * N * FLOP_PER_PT
    print(f"ErrorEstimator B={B} N={N}: cuDNN fp32 {t32:.3f} ms ({fl/t32/1e9:.1f} TFLOP/s) | tcgen05 bf16 {ttc:.3f} ms ({fl/ttc/1e9:.1f} TFLOP/s) | x{t32/ttc:.1f}", flush=True)
    out[f"mlp_B{B}"] = {"fp32_ms": t32, "tc_ms": ttc}
    # GEMM kernel alone, the three fat layers
    lib = _lib.lib()
    Npad = (N + 127) // 128 * 128
    for K, Co in [(128, 1024), (1024, 512), (512, 256)]:
        X = torch.randn(B * Npad, K, device="cuda").bfloat16(); W = torch.randn(Co, K, device="cuda").bfloat16()
        bias = torch.zeros(Co, device="cuda"); Y = torch.empty(B * Npad, Co, device="cuda", dtype=torch.bfloat16)
        stats = torch.zeros(B, Co, 2, device="cuda")
        t = ev(lambda: lib.fepe_mlp_gemm(X.data_ptr(), W.data_ptr(), bias.data_ptr(), Y.data_ptr(), stats.data_ptr(), B, Npad, N, K, Co, torch.cuda.current_stream().cuda_stream))
        print(f"   gemm K={K} Co={Co}: {t*1e3:.1f} us  {2*B*Npad*K*Co/t/1e9:.1f} TFLOP/s  ({(B*Npad*(K+Co)*2)/t/1e6:.0f} GB/s activations)", flush=True)
for B, N in [(64, 1000), (512, 1000)]:
    net = DeepFNet(depth=5, image_size=[376, 1241, 3], if_quality=False).cuda()
    d = synth.make_batch(min(B, 64), N, seed=1)
    m = torch.from_numpy(d["matches_xy_ori"]).cuda().repeat((B + 63) // 64, 1, 1)[:B].contiguous()
    batch = {"matches_xy_ori": m}
    with torch.no_grad():
        t32 = ev(lambda: net(batch), iters=5, warm=2)
        net.enable_tensor_core_mlp()
        ttc = ev(lambda: net(batch), iters=5, warm=2)
    print(f"DeepFNet forward depth 5 B={B} N={N}: fp32 MLP {t32:.2f} ms ({B/t32*1e3:.0f} pairs/s) | tcgen05 MLP {ttc:.2f} ms ({B/ttc*1e3:.0f} pairs/s)", flush=True)
    out[f"deepf_B{B}"] = {"fp32_ms": t32, "tc_ms": ttc}
print(json.dumps(out))

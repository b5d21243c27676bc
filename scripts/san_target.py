"""compute-sanitizer target: every kernel once on small shapes (ragged and aligned N)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import torch
from fepe_b200 import ops, synth
from fepe_b200 import _lib as _fepe_lib
from fepe_b200.models import ErrorEstimator
shapes = [(5, 333), (3, 37), (4, 1000)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for kern in ("ring", "small"):
    _fepe_lib.set_dispatch("fit", kern)
    for B, N in shapes:
        d = synth.make_batch(B, N, seed=1)
        aff = ops.hw_affine(d["image_size"])
        m = torch.from_numpy(d["matches_xy_ori"]).cuda()
        w = torch.from_numpy(d["weights"]).cuda().requires_grad_(True)
        F, r, e = ops.FitFunction.apply(m, w.reshape(B, N), *aff, 0.5)
        (F.sum() + r.sum() + e.sum()).backward()
        mc = m.clone().requires_grad_(True)              # coordinate-gradient path (fepe_fit_bwd_coords)
        F2, r2, e2 = ops.FitFunction.apply(mc, w.reshape(B, N), *aff, 0.5)
        (F2.sum() + r2.sum() + e2.sum()).backward()
        t = lambda k: torch.from_numpy(d[k]).cuda()
        ops.pose_forward(F.detach(), t("Ks"), aff, t("q_cam"), t("t_cam"), t("delta_Rtijs_4_4"), t("pts1_virt"), t("pts2_virt"))
ee = ErrorEstimator(4).cuda()
ee.tensor_cores = True
with torch.no_grad():
    ee(torch.rand(2, 4, 333, device="cuda"))
torch.cuda.synchronize()
print("sanitizer target done")
d = synth.make_batch(3, 333, seed=2)
t = lambda k: torch.from_numpy(d[k]).cuda()
ops.recover_pose(t("E_gt"), t("Ks"), t("matches_xy_ori"), t("delta_Rtijs_4_4"), want_mask=True)
torch.cuda.synchronize()
print("recover_pose done")
rng = torch.Generator(device="cuda").manual_seed(0)
d1 = torch.nn.functional.normalize(torch.randn(2, 333, 256, device="cuda", generator=rng), dim=2)
d2 = torch.nn.functional.normalize(torch.randn(2, 200, 256, device="cuda", generator=rng), dim=2)
ops.nn_match_two_way(d1, d2, 1.0, torch.tensor([333, 100], dtype=torch.int32, device="cuda"), None)
torch.cuda.synchronize()
print("nn_match done")

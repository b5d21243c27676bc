"""The reference's OWN callers executed on the product: `utils/loader.py:modelLoader` builds the model through the
integration shim, and `train_good_utils.get_all_loss_DeepF` / `get_Rt_loss` (unmodified, imported from oracle/_ref)
consume the dict `DeepFNet.forward` returns -- the "drops in unchanged" claim of INTEGRATION.md as a test.
The reference sources come from oracle/_ref (oracle/make_ref.py; present on the GPU box) or /root/reference."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import fepe_oracle as O
from oracle import ref_env

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = torch.from_numpy
# the kwargs deepFEPE/train_good.py:176-189 passes (configs/kitti_corr_baseline.yaml)
MODEL_KW = dict(depth=5, img_zoom_xy=[1.0, 1.0], image_size=[376, 1241, 3], quality_size=0, if_quality=False,
                if_img_des_to_pointnet=False, if_goodCorresArch=False, if_img_feat=False, if_cpu_svd=True,
                if_learn_offsets=False, if_tri_depth=False, if_sample_loss=False)

needs_ref = pytest.mark.skipif(ref_env.reference_root() is None, reason="oracle/_ref not built (python oracle/make_ref.py)")


def _loader():
    ref = ref_env.import_reference()
    shim = os.path.join(ROOT, "integration")
    if shim not in sys.path or sys.path.index(shim) > sys.path.index(os.path.join(ref.root, "deepFEPE")):
        sys.path.insert(0, shim)                      # `models` must resolve to the shim, not to deepFEPE/models
    assert "models" not in sys.modules or sys.modules["models"].__file__.startswith(shim)
    with ref_env.quiet():
        import utils.loader as loader                 # the reference's deepFEPE/utils/loader.py
    assert loader.__file__.startswith(ref.root)
    return ref, loader


@needs_ref
def test_reference_model_loader_builds_the_product_model(golden):
    """deepFEPE/utils/loader.py:117-129 unmodified: model.name -> `from models.DeepFNet import DeepFNet` -> our class,
    with the kwargs train_good.py passes and the reference's state_dict keys."""
    ref, loader = _loader()
    net = loader.modelLoader("GoodCorresNet_layers_deepF", **MODEL_KW)
    from fepe_b200.models import DeepFNet as OurDeepFNet
    assert type(net) is OurDeepFNet
    assert sorted(net.state_dict().keys()) == sorted(golden["c1_state_keys"].tolist())


@needs_ref
@pytest.mark.gpu
def test_reference_loss_glue_runs_on_product_outputs():
    """outs = product DeepFNet(data_batch); the reference's get_all_loss_DeepF (train_good_utils.py:298-520, incl.
    loss_epi_res :429-438, loss_min_batch, the residual / weight regularisers) and get_Rt_loss (:64-295) run on it
    unchanged; their numbers agree with the product's device loss head, and backward reaches both MLPs."""
    from fepe_b200 import ops, synth
    from fepe_b200 import losses as L
    ref, loader = _loader()
    torch.manual_seed(3)
    net = loader.modelLoader("GoodCorresNet_layers_deepF", **MODEL_KW).cuda()
    B, N = 4, 512
    d = synth.make_batch(B, N, seed=17)
    c = lambda k: T(d[k]).cuda()
    data_batch = {"matches_xy_ori": c("matches_xy_ori"), "matches_good_unique_nums": T(d["matches_good_unique_nums"]),
                  "t_scene_scale": torch.ones(B, 1, 1).cuda(), "Ks": c("Ks"), "K_invs": c("K_invs")}
    outs = net(data_batch)
    loss_params = {"depth": 5, "clamp_at": 0.02, "if_tri_depth": False, "if_sample_loss": False, "topK": 20,
                   "matches_good_unique_nums": d["matches_good_unique_nums"].tolist(), "model": "GoodCorresNet_layers_deepF"}
    with ref_env.quiet():
        losses, E_ests, F_ests, w_soft, rn, rnmax, E_layers = ref.tgu.get_all_loss_DeepF(
            outs, c("pts1_virt"), c("pts2_virt"), c("Ks"), loss_params, get_residual_summaries=True)
    for k in ("loss_F", "loss_epi_res", "loss_residual", "loss_regW_entro", "loss_min_batch", "loss_min_layers"):
        assert torch.isfinite(torch.as_tensor(losses[k])).all(), k
    assert len(losses["loss_layers"]) == 5 and len(losses["loss_epi_res_layers"]) == 4 and len(E_layers) == 5
    # the product's device head (fepe_pose_fwd) computes the same F-loss and E per layer
    Fl = torch.stack([o.detach() for o in outs["out_layers"]])
    aff = ops.hw_affine([376, 1241])
    pose = ops.pose_forward(Fl, c("Ks"), aff, c("q_cam"), c("t_cam"), c("delta_Rtijs_4_4"), c("pts1_virt"), c("pts2_virt"), 0.02)
    for l in range(5):
        assert abs(float(pose[l, :, 25].mean()) - float(losses["loss_layers"][l])) < 2e-5
        Er = E_layers[l].detach()
        assert float((pose[l, :, :9].reshape(B, 3, 3) - Er).abs().max() / Er.abs().max()) < 1e-4
    # the reference's pose loss (host loop with LAPACK 3x3 SVDs) against the product's device mirror of the same name
    with ref_env.quiet():
        rt_ref = ref.tgu.get_Rt_loss(E_layers, T(d["Ks"]), T(d["matches_xy_ori"][:, :, :2]), T(d["matches_xy_ori"][:, :, 2:]),
                                     T(d["delta_Rtijs_4_4"]), c("q_cam"), c("t_cam"), device="cuda")
    rt_ours = L.get_Rt_loss(E_layers, None, None, None, c("delta_Rtijs_4_4"), c("q_cam"), c("t_cam"))
    for key in ("q_l2_error_layers_list", "t_l2_error_layers_list"):
        a = torch.stack([x.reshape(-1) for x in rt_ref[key]]).detach().cpu()
        b = torch.stack(rt_ours[key]).detach().cpu()
        assert float((a - b).abs().max()) < 5e-5, key
    for key in ("R_angle_error_layers_list", "t_angle_error_layers_list"):
        a, b = np.stack(rt_ref[key]).reshape(5, B), np.stack(rt_ours[key])
        assert np.abs(a - b).max() < 2e-2, key                      # degrees; the reference's angles come from fp32 matrices
    # a training step exactly as Train_model_pipeline.py:560-595 builds it, through the reference's loss functions
    q = torch.stack([x.reshape(-1) for x in rt_ref["q_l2_error_layers_list"]])
    t = torch.stack([x.reshape(-1) for x in rt_ref["t_l2_error_layers_list"]])
    loss = losses["loss_F"] + torch.clamp(q, 0, 0.1).mean() * 1.0 + torch.clamp(t, 0, 0.5).mean() * 0.1
    loss.backward()
    for name, prm in net.named_parameters():
        assert prm.grad is not None and torch.isfinite(prm.grad).all(), name
    assert float(net.input_weights.fw[0].weight.grad.abs().sum()) > 0
    assert float(net.update_weights.fw[0].weight.grad.abs().sum()) > 0

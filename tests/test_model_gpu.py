"""GPU tests of the drop-in module surface: DeepFNet / Fit with the reference's constructor, dict keys
and state_dict keys, forward vs the oracle's restatement of DeepFNet.forward, and gradients reaching
both MLPs."""
import numpy as np
import pytest
import torch

from oracle import fepe_oracle as O
from fepe_b200 import synth
from fepe_b200.models import DeepFNet, Fit, ErrorEstimator

pytestmark = pytest.mark.gpu
T = torch.from_numpy
MODEL_KW = dict(depth=5, image_size=[376, 1241, 3], quality_size=0, if_quality=False, if_img_des_to_pointnet=False,
                if_goodCorresArch=False, if_img_feat=False, if_cpu_svd=True, if_learn_offsets=False,
                if_tri_depth=False, if_sample_loss=False)   # what train_good.py:176-189 passes


def _batch(d):
    return {"matches_xy_ori": T(d["matches_xy_ori"]).cuda(),
            "matches_good_unique_nums": T(d["matches_good_unique_nums"]),
            "t_scene_scale": torch.ones(d["matches_xy_ori"].shape[0], 1, 1).cuda(),
            "Ks": T(d["Ks"]).cuda(), "K_invs": T(d["K_invs"]).cuda()}


def test_state_dict_keys_match_reference(golden):
    net = DeepFNet(**MODEL_KW)
    assert sorted(net.state_dict().keys()) == sorted(golden["c1_state_keys"].tolist())


def test_forward_matches_oracle_and_reference_layer0(golden):
    torch.manual_seed(77)
    net = DeepFNet(**MODEL_KW).cuda()       # same construction order / seed as the golden generator
    o_init, o_upd = O.build_error_estimator(4), O.build_error_estimator(7)
    o_init.load_state_dict(net.input_weights.fw.state_dict())
    o_upd.load_state_dict(net.update_weights.fw.state_dict())
    m = T(golden["c1b_matches"])
    with torch.no_grad():
        outs = net({"matches_xy_ori": m.cuda(), "matches_good_unique_nums": torch.tensor([160, 160]),
                    "t_scene_scale": torch.ones(2, 1, 1).cuda()})
        ref = O.deepf_forward(m, [376, 1241, 3], o_init.cpu(), o_upd.cpu(), depth=5)
    for key in ("logits", "logits_layers", "F_est", "epi_res_layers", "T1", "T2", "out_layers", "pts1", "pts2",
                "weights", "residual_layers", "weights_layers"):
        assert key in outs
    assert len(outs["out_layers"]) == 5 and len(outs["epi_res_layers"]) == 4
    # layer 0 is independent of the arbitrary sign of f: compare with the reference's own output
    # (round 2: the weight network runs on the fp32-parity tensor-core path; measured 1.4e-5 on the weights, 1.6e-6 on F.
    # All five layers are compared in tests/test_all_layers.py.)
    np.testing.assert_allclose(outs["weights_layers"][0].cpu().numpy(), golden["c1b_w_layers"][0], rtol=3e-4, atol=1e-8)
    err0 = O.sign_aligned_rel_err(outs["out_layers"][0].cpu(), T(golden["c1b_F_layers"][0]))
    assert float(err0.max()) < 1e-4
    np.testing.assert_allclose(outs["epi_res_layers"][0].cpu().numpy(), golden["c1b_epi_layers"][0], atol=2e-4)
    assert outs["pts1"].shape == (2, 160, 3) and outs["T1"].shape == (2, 3, 3)
    np.testing.assert_allclose(outs["pts1"].cpu().numpy(), ref["pts1"].numpy(), atol=1e-6)
    for l in range(5):
        assert torch.isfinite(outs["out_layers"][l]).all()
        # rank 2 at every layer
        assert float(torch.linalg.svdvals(outs["out_layers"][l].double())[:, 2].max()) < 1e-6


def test_training_step_reaches_both_mlps():
    torch.manual_seed(0)
    net = DeepFNet(**MODEL_KW).cuda()
    d = synth.make_batch(4, 512, seed=5)
    outs = net(_batch(d))
    # F-loss exactly as get_all_loss_DeepF builds it (train_good_utils.py:325-364) with torch ops
    T1 = outs["T1"]
    p1 = (T1 @ T(d["pts1_virt"]).cuda().permute(0, 2, 1)).permute(0, 2, 1)
    p2 = (T1 @ T(d["pts2_virt"]).cuda().permute(0, 2, 1)).permute(0, 2, 1)
    loss = sum(O.epi_residual(p1, p2, Fo, 0.02).mean() for Fo in outs["out_layers"]) / 5
    loss.backward()
    for name, prm in net.named_parameters():
        assert prm.grad is not None and torch.isfinite(prm.grad).all(), name
    assert float(net.input_weights.fw[0].weight.grad.abs().sum()) > 0
    assert float(net.update_weights.fw[0].weight.grad.abs().sum()) > 0


def test_fit_module_signature():
    d = synth.make_batch(3, 200, seed=2)
    p1, p2, _ = O.norm_hw(T(d["matches_xy_ori"]), d["image_size"])
    fit = Fit(is_cuda=True, is_test=False, if_cpu_svd=True)
    out, res = fit(p1.cuda(), p2.cuda(), T(d["weights"]).cuda())
    Fr, rr = O.fit_weighted_svd(p1, p2, T(d["weights"]))
    assert out.shape == (3, 3, 3) and res.shape == (3, 200)
    assert float(O.sign_aligned_rel_err(out.cpu(), Fr).max()) < 1e-4


def test_error_estimator_matches_reference_output(golden):
    torch.manual_seed(1234)
    ee = ErrorEstimator(4).cuda()
    with torch.no_grad():
        y = ee(T(golden["ee_x"]).cuda())
    np.testing.assert_allclose(y.cpu().numpy(), golden["ee_y"], rtol=1e-4, atol=2e-5)


def test_learn_offsets_matches_reference_forward_and_backward():
    """DeepFNet(if_learn_offsets=True) against the UNMODIFIED reference module run in fp64 on the CPU
    (tests/golden/make_golden_offsets.py -> reference_offsets.npz): same seed => same initial parameters; offsets,
    per-layer F and the parameter gradients of all three networks (the offsets net only gets a gradient through the
    coordinate gradient of the fit, fepe_fit_bwd_coords).  Tolerances: the networks run in fp32 here vs fp64 there."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_offsets.npz"), allow_pickle=False))
    torch.manual_seed(78)
    net = DeepFNet(depth=3, image_size=[376, 1241, 3], if_quality=False, is_cuda=True, if_cpu_svd=False,
                   if_learn_offsets=True).cuda()
    assert sorted(net.state_dict().keys()) == sorted(g["state_keys"].tolist())
    m = T(g["matches"]).cuda()
    outs = net({"matches_xy_ori": m, "matches_good_unique_nums": torch.tensor([200, 200]),
                "t_scene_scale": torch.ones(2, 1, 1).cuda()})
    assert outs["offsets"].shape == (2, 4, 200)
    np.testing.assert_allclose(outs["offsets"].detach().cpu().numpy(), g["offsets"], atol=5e-3)
    np.testing.assert_allclose(outs["pts1"].detach().cpu().numpy(), g["pts1"], atol=1e-5)
    for l in range(3):
        err = O.sign_aligned_rel_err(outs["out_layers"][l].detach().cpu(), T(g["F_layers"][l]))
        print("layer", l, "F rel err", err.tolist())
        assert float(err.max()) < 5e-3
    loss = 0.0
    for Fl in outs["out_layers"]:
        loss = loss + O.epi_residual(outs["pts1"], outs["pts2"], Fl, 0.1).mean()
    for r in outs["residual_layers"]:
        loss = loss + 1e3 * (r ** 2).sum(1).mean()
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-3 * abs(float(g["loss"]))
    params = dict(net.named_parameters())
    worst = 0.0
    for key, ref in g.items():
        if not key.startswith("grad/") or np.linalg.norm(ref) < 1e-8:      # conv biases in front of an InstanceNorm: 0
            continue
        got = params[key[5:]].grad.detach().cpu().numpy().astype(np.float64)
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(key, "rel grad err %.2e" % rel)
        worst = max(worst, rel)
    assert float(params["update_offsets.fw.15.weight"].grad.abs().sum()) > 0
    assert worst < 5e-2, worst


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_replicas_match_single_device():
    """The reference's multi-GPU mode (nn.DataParallel, deepFEPE/train_good.py:309-314): the dict input is scattered along
    dim 0, the module replicated; every replica must launch on its own device / stream and use its own weights -- in
    inference and under autograd."""
    torch.manual_seed(5)
    net = DeepFNet(**MODEL_KW).cuda(0)
    d = synth.make_batch(4, 512, seed=12)
    batch = {"matches_xy_ori": T(d["matches_xy_ori"]).cuda(0), "matches_good_unique_nums": T(d["matches_good_unique_nums"]).cuda(0),
             "t_scene_scale": torch.ones(4, 1, 1).cuda(0)}
    with torch.no_grad():
        single = net(batch)
        dp = torch.nn.DataParallel(net, device_ids=[0, 1])
        multi = dp(batch)
        again = dp(batch)                                       # replicas are rebuilt every call
    for a, b, c in zip(single["out_layers"], multi["out_layers"], again["out_layers"]):
        assert b.device.index == 0 and b.shape == a.shape
        assert float(O.sign_aligned_rel_err(b.cpu(), a.cpu()).max()) < 1e-5
        assert float(O.sign_aligned_rel_err(c.cpu(), a.cpu()).max()) < 1e-5
    outs = dp(batch)
    sum(Fo.pow(2).sum() for Fo in outs["out_layers"]).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())

"""Pin the CPU oracle (oracle/fepe_oracle.py) to outputs of the UNMODIFIED reference
(tests/golden/reference_outputs.npz, produced by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import fepe_oracle as O
from fepe_b200 import synth

T = torch.from_numpy
N_FIT_CASES = 7


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("i", range(N_FIT_CASES))
def test_fit_matches_reference(golden, i):
    B, N, seed = golden[f"fit{i}_meta"]
    mode = str(golden[f"fit{i}_mode"])
    d = synth.make_batch(int(B), int(N), int(seed), weight_mode=mode)
    p1, p2, Tn = O.norm_hw(T(d["matches_xy_ori"]), d["image_size"])
    np.testing.assert_allclose(p1.numpy(), golden[f"fit{i}_pts1"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(Tn.numpy(), golden[f"fit{i}_T1"], rtol=0, atol=0)
    # use the reference's own normalised points so only Fit is under test from here on
    p1, p2 = T(golden[f"fit{i}_pts1"]), T(golden[f"fit{i}_pts2"])
    Fo, res = O.fit_weighted_svd(p1, p2, T(d["weights"]))
    Fr = T(golden[f"fit{i}_F"])
    err = O.sign_aligned_rel_err(Fo, Fr)
    assert float(err.max()) < 2e-5, err
    sgn = torch.sign((Fo * Fr).sum((1, 2))).view(-1, 1)
    np.testing.assert_allclose((res * sgn).numpy(), golden[f"fit{i}_res"], rtol=0, atol=2e-6)
    # with the reference's F the residual / loss / E restatements must agree tightly
    epi = O.epi_residual(p1, p2, Fr)
    np.testing.assert_allclose(epi.numpy(), golden[f"fit{i}_epi"], rtol=1e-5, atol=1e-6)
    lossF, losses, E_layers = O.f_loss_layers([Fr], Tn, Tn, T(d["pts1_virt"]), T(d["pts2_virt"]),
                                              T(d["Ks"]), clamp_at=0.02)
    np.testing.assert_allclose(losses[0].numpy(), golden[f"fit{i}_lossF"], rtol=1e-5, atol=1e-7)
    assert _rel(E_layers[0].numpy(), golden[f"fit{i}_E"]) < 1e-6


def test_fit_backward_matches_reference_autograd(golden):
    p1, p2 = T(golden["bwd_pts1"]), T(golden["bwd_pts2"])
    w = T(golden["bwd_w"]).clone().requires_grad_(True)
    Fo, res = O.fit_weighted_svd(p1, p2, w)
    epi = O.epi_residual(p1, p2, Fo)
    sgn = torch.sign(Fo.detach()[:, 2, 2]).view(-1, 1, 1)
    loss = ((Fo * sgn) * T(golden["bwd_gF"])).sum() + ((res * sgn.view(-1, 1)) * T(golden["bwd_gr"])).sum() \
        + (epi * T(golden["bwd_ge"])).sum()
    loss.backward()
    assert _rel(w.grad.numpy(), golden["bwd_gw"]) < 1e-8
    assert float(O.sign_aligned_rel_err(Fo.detach(), T(golden["bwd_F"])).max()) < 1e-10


def test_pose_matches_reference(golden):
    E = T(golden["pose_E"])
    for b in range(E.shape[0]):
        Rs, ts = O.essential_decompose(E[b].t())
        np.testing.assert_allclose(Rs[0].numpy(), golden["pose_R1"][b], atol=1e-6)
        np.testing.assert_allclose(Rs[1].numpy(), golden["pose_R2"][b], atol=1e-6)
        np.testing.assert_allclose(ts[0].numpy(), golden["pose_t"][b], atol=1e-6)
        np.testing.assert_allclose(O.rot_to_quat(Rs[0]).numpy(), golden["pose_q1"][b], atol=1e-6)
        np.testing.assert_allclose(O.rot_to_quat(Rs[1]).numpy(), golden["pose_q2"][b], atol=1e-6)
    q, t, ra, ta = O.pose_errors(E, T(golden["pose_qcam"]), T(golden["pose_tcam"]), T(golden["pose_Rt"]))
    np.testing.assert_allclose(q.numpy(), golden["pose_q_l2"][0], atol=1e-6)
    np.testing.assert_allclose(t.numpy(), golden["pose_t_l2"][0], atol=1e-6)
    np.testing.assert_allclose(ra.numpy(), golden["pose_R_ang"][0], atol=2e-3)   # cv2.Rodrigues vs acos
    np.testing.assert_allclose(ta.numpy(), golden["pose_t_ang"][0], atol=2e-3)
    # sanity on the synthetic scene: the estimated pose is close to the GT one
    assert float(ra.max()) < 1.0 and float(ta.max()) < 10.0


def test_quaternion_branches(golden):
    for R, q in zip(golden["quat_R"], golden["quat_q"]):
        np.testing.assert_allclose(O.rot_to_quat(T(R)).numpy(), q, atol=1e-6)
        np.testing.assert_allclose(synth.rot_to_quat(R)[:, None], q, atol=1e-6)


def test_error_estimator_matches_reference(golden):
    torch.manual_seed(1234)
    ee = O.build_error_estimator(4)
    assert list(ee.state_dict().keys()) == [k[3:] if k.startswith("fw.") else k for k in golden["ee_keys"]]
    with torch.no_grad():
        y = ee(T(golden["ee_x"]))
    np.testing.assert_allclose(y.numpy(), golden["ee_y"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag", ["c1", "c1b"])
def test_deepf_forward_matches_reference(golden, tag):
    """Config 1 (plumbing): 1 pair x 100 planar correspondences through the whole DeepFNet
    loop, and a non-degenerate 2x160 batch.  Planar scenes make F non-unique, so the
    planar case checks weights/logits/epipolar residuals, the generic one also F."""
    torch.manual_seed(77)
    net_init = O.build_error_estimator(4)
    net_upd = O.build_error_estimator(7)
    with torch.no_grad():
        o = O.deepf_forward(T(golden[f"{tag}_matches"]), [376, 1241, 3], net_init, net_upd, depth=5)
    w_ref = golden[f"{tag}_w_layers"]
    np.testing.assert_allclose(o["weights_layers"][0].numpy(), w_ref[0], rtol=1e-4, atol=1e-7)
    if tag == "c1b":
        for l in range(5):
            err = O.sign_aligned_rel_err(o["out_layers"][l], T(golden[f"{tag}_F_layers"][l]))
            assert float(err.max()) < 1e-3, (l, err)
            np.testing.assert_allclose(o["weights_layers"][l].numpy(), w_ref[l], rtol=2e-2, atol=1e-5)
        np.testing.assert_allclose(torch.stack(o["epi_res_layers"]).numpy(), golden[f"{tag}_epi_layers"],
                                   rtol=0, atol=2e-3)
    else:
        assert all(torch.isfinite(x).all() for x in o["out_layers"])
        assert torch.stack(o["weights_layers"]).shape == tuple(w_ref.shape)


def test_state_dict_keys_of_reference_deepfnet(golden):
    keys = set(golden["c1_state_keys"].tolist())
    assert "input_weights.fw.0.weight" in keys and "update_weights.fw.15.bias" in keys

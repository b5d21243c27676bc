"""GPU parity of the pose / loss head (fepe_pose_fwd) against the oracle and the reference's
get_Rt_loss outputs.  Tolerances: E 1e-5 relative; q/t L2 errors 2e-5 absolute; angles 1e-3 deg
(SURVEY 8c); F-loss 1e-6 absolute."""
import numpy as np
import pytest
import torch

from oracle import fepe_oracle as O
from fepe_b200 import ops, synth

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def test_against_reference_golden(golden):
    """Feed the reference's own E back through F = (TK)^-T E (TK)^-1 is awkward; instead check the
    decomposition outputs for the golden E by passing K = I, T = I (so E == F)."""
    E = T(golden["pose_E"]).cuda()
    B = E.shape[0]
    I = torch.eye(3, device="cuda").expand(B, 3, 3).contiguous()
    out = ops.pose_forward(E, I, ops.IDENTITY_AFFINE, T(golden["pose_qcam"]).cuda(), T(golden["pose_tcam"]).cuda(),
                           T(golden["pose_Rt"]).cuda())[0].cpu().numpy()
    np.testing.assert_allclose(out[:, 21], golden["pose_q_l2"][0], atol=2e-5)
    np.testing.assert_allclose(out[:, 22], golden["pose_t_l2"][0], atol=2e-5)
    np.testing.assert_allclose(out[:, 23], golden["pose_R_ang"][0], atol=2e-3)
    np.testing.assert_allclose(out[:, 24], golden["pose_t_ang"][0], atol=2e-3)


@pytest.mark.parametrize("B,N,L", [(64, 2000, 1), (32, 500, 3)])
def test_against_oracle(B, N, L):
    d = synth.make_batch(B, N, seed=40 + L, weight_mode="inlier", outlier_frac=0.2)
    aff = ops.hw_affine(d["image_size"])
    m, w = T(d["matches_xy_ori"]).cuda(), T(d["weights"]).cuda()
    Fs = []
    for l in range(L):
        F, _, _, _ = ops.fit_forward(m, w * (1.0 + 0.1 * l) if l else w, aff)
        # perturb later layers a little so that layers differ
        Fs.append(F + 1e-3 * l * torch.randn(F.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(l)))
    Fl = torch.stack(Fs)
    out = ops.pose_forward(Fl, T(d["Ks"]).cuda(), aff, T(d["q_cam"]).cuda(), T(d["t_cam"]).cuda(),
                           T(d["delta_Rtijs_4_4"]).cuda(), T(d["pts1_virt"]).cuda(), T(d["pts2_virt"]).cuda(),
                           clamp_at=0.02).cpu()
    _, _, Tn = O.norm_hw(T(d["matches_xy_ori"]), d["image_size"])
    lossF, losses, E_layers = O.f_loss_layers([f.cpu() for f in Fl], Tn, Tn, T(d["pts1_virt"]), T(d["pts2_virt"]),
                                              T(d["Ks"]), clamp_at=0.02)
    for l in range(L):
        E = E_layers[l]
        rel = (out[l, :, :9].reshape(B, 3, 3) - E).flatten(1).norm(dim=1) / E.flatten(1).norm(dim=1)
        assert float(rel.max()) < 1e-5
        q, t, ra, ta = O.pose_errors(E.double(), T(d["q_cam"]).double(), T(d["t_cam"]).double(),
                                     T(d["delta_Rtijs_4_4"]).double())
        np.testing.assert_allclose(out[l, :, 21].numpy(), q.numpy(), atol=2e-5)
        np.testing.assert_allclose(out[l, :, 22].numpy(), t.numpy(), atol=2e-5)
        np.testing.assert_allclose(out[l, :, 23].numpy(), ra.numpy(), atol=1e-3)
        np.testing.assert_allclose(out[l, :, 24].numpy(), ta.numpy(), atol=1e-3)
        np.testing.assert_allclose(out[l, :, 25].numpy(), losses[l].mean(1).numpy(), atol=1e-6, rtol=1e-4)
        Rsel = out[l, :, 9:18].reshape(B, 3, 3).double()
        assert float((Rsel @ Rsel.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max()) < 1e-5
    # on this synthetic scene (20 % outliers, inlier favouring weights) the pose is recovered
    assert float(out[0, :, 23].median()) < 0.5 and float(out[0, :, 24].median()) < 5.0


def test_gt_fundamental_gives_zero_pose_error():
    d = synth.make_batch(16, 8, seed=3)
    aff = ops.hw_affine(d["image_size"])
    # F_gt is in pixel coordinates; move it to primed coordinates: F' = T^-T F T^-1
    Tn = synth.norm_hw_transform(d["image_size"])
    Ti = np.linalg.inv(Tn)
    Fp = (Ti.T @ d["F_gt"].astype(np.float64) @ Ti).astype(np.float32)
    out = ops.pose_forward(T(Fp).cuda(), T(d["Ks"]).cuda(), aff, T(d["q_cam"]).cuda(), T(d["t_cam"]).cuda(),
                           T(d["delta_Rtijs_4_4"]).cuda(), T(d["pts1_virt"]).cuda(), T(d["pts2_virt"]).cuda())[0].cpu()
    assert float(out[:, 21].max()) < 1e-4 and float(out[:, 22].max()) < 1e-3
    assert float(out[:, 23].max()) < 1e-2 and float(out[:, 24].max()) < 0.1
    assert float(out[:, 25].max()) < 1e-4      # "SHOULD BE ALL ZEROS" (utils_misc.py:174)


def test_pose_loss_function_gradient_against_oracle_autograd():
    """dL/dF of the device head vs torch autograd (fp64) through the oracle's F-loss + pose errors."""
    B, L = 16, 2
    d = synth.make_batch(B, 400, seed=11, weight_mode="inlier", outlier_frac=0.1)
    aff = ops.hw_affine(d["image_size"])
    F0, _, _, _ = ops.fit_forward(T(d["matches_xy_ori"]).cuda(), T(d["weights"]).cuda(), aff)
    gen = torch.Generator("cuda").manual_seed(0)
    Fl = torch.stack([F0, F0 + 2e-3 * torch.randn(F0.shape, device="cuda", generator=gen)]).requires_grad_(True)
    cuda = lambda k: T(d[k]).cuda()
    q, t, lf, out = ops.PoseLossFunction.apply(Fl, cuda("Ks"), cuda("q_cam"), cuda("t_cam"), cuda("delta_Rtijs_4_4"),
                                                cuda("pts1_virt"), cuda("pts2_virt"), *aff, 0.02)
    gq = torch.randn(L, B, device="cuda", generator=gen)
    gt = torch.randn(L, B, device="cuda", generator=gen)
    gl = torch.randn(L, B, device="cuda", generator=gen)
    ((q * gq).sum() + (t * gt).sum() + (lf * gl).sum()).backward()
    ours = Fl.grad.cpu().double()
    # oracle in fp64
    Fr = Fl.detach().cpu().double().requires_grad_(True)
    _, _, Tn = O.norm_hw(T(d["matches_xy_ori"]).double(), d["image_size"])
    loss = 0.0
    for l in range(L):
        _, losses, E_layers = O.f_loss_layers([Fr[l]], Tn, Tn, T(d["pts1_virt"]).double(), T(d["pts2_virt"]).double(),
                                              T(d["Ks"]).double(), clamp_at=0.02)
        ql, tl, _, _ = O.pose_errors(E_layers[0], T(d["q_cam"]).double(), T(d["t_cam"]).double(),
                                     T(d["delta_Rtijs_4_4"]).double())
        loss = loss + (ql * gq[l].cpu().double()).sum() + (tl * gt[l].cpu().double()).sum() \
            + (losses[0].mean(1) * gl[l].cpu().double()).sum()
    loss.backward()
    ref = Fr.grad
    rel = (ours - ref).flatten(2).norm(dim=2) / ref.flatten(2).norm(dim=2)
    print("pose head dL/dF rel err: median %.2e max %.2e" % (float(rel.median()), float(rel.max())))
    # fp32 F and fp32 virtual-point arithmetic on our side; SVD-adjoint denominators 1/(s1^2-s2^2) amplify that
    assert float(rel.median()) < 1e-3 and float(rel.max()) < 5e-2


@pytest.mark.parametrize("B,N,with_virt,with_rt", [(256, 1000, True, True), (40, 333, True, False), (7, 64, False, True),
                                                   (600, 200, True, True)])
def test_fused_fit_pose_equals_two_calls(B, N, with_virt, with_rt):
    """fepe_fit_pose_fwd (pose head inside the latency kernel for B <= 2 pairs per SM, two launches above) against
    fepe_fit_fwd + fepe_pose_fwd on the same inputs."""
    d = synth.make_batch(B, N, seed=91 + N, weight_mode="softmax", outlier_frac=0.3)
    aff = ops.hw_affine(d["image_size"])
    g = lambda k: T(d[k]).cuda()
    m, w = g("matches_xy_ori"), g("weights")
    v1, v2 = (g("pts1_virt"), g("pts2_virt")) if with_virt else (None, None)
    rt = g("delta_Rtijs_4_4") if with_rt else None
    F0, r0, e0, s0 = ops.fit_forward(m, w, aff, want_saved=True)
    p0 = ops.pose_forward(F0, g("Ks"), aff, g("q_cam"), g("t_cam"), rt, v1, v2, clamp_at=0.02)[0]
    F1, r1, e1, s1, p1 = ops.fit_pose_forward(m, w, aff, g("Ks"), g("q_cam"), g("t_cam"), rt, v1, v2,
                                              virt_clamp_at=0.02, want_saved=True)
    torch.cuda.synchronize()
    assert torch.equal(F0, F1) and torch.equal(r0, r1) and torch.equal(e0, e1)
    assert torch.equal(s0[:, :56], s1[:, :56])
    np.testing.assert_allclose(p1.cpu().numpy(), p0.cpu().numpy(), rtol=1e-6, atol=1e-6)


def test_get_Rt_loss_mirror_matches_oracle_values_and_gradient(golden):
    """fepe_b200.losses.get_Rt_loss (the reference's name / arguments / dict, train_good_utils.py:64-295) against the
    oracle's restatement of the same loop in fp64, values and d loss / d E, and against the reference's own numbers."""
    from fepe_b200.losses import get_Rt_loss, pose_loss_from_Rt_loss
    # (1) the reference's own outputs for its own E (golden fixture)
    E = T(golden["pose_E"]).cuda()
    B = E.shape[0]
    r = get_Rt_loss([E], None, None, None, T(golden["pose_Rt"]), T(golden["pose_qcam"]), T(golden["pose_tcam"]), device="cuda")
    np.testing.assert_allclose(r["q_l2_error_layers_list"][0].cpu().numpy(), golden["pose_q_l2"][0], atol=2e-5)
    np.testing.assert_allclose(r["t_l2_error_layers_list"][0].cpu().numpy(), golden["pose_t_l2"][0], atol=2e-5)
    np.testing.assert_allclose(r["R_angle_error_layers_list"][0], golden["pose_R_ang"][0], atol=2e-3)
    np.testing.assert_allclose(r["t_angle_error_layers_list"][0], golden["pose_t_ang"][0], atol=2e-3)
    assert set(r) == {"t_l2_error_mean", "q_l2_error_mean", "t_l2_error_list", "q_l2_error_list", "R_angle_error_mean",
                      "R_angle_error_list", "t_angle_error_mean", "t_angle_error_list", "R_angle_error_layers_list",
                      "t_angle_error_layers_list", "t_l2_error_layers_list", "q_l2_error_layers_list"}
    # (2) three layers of perturbed essential matrices: loss value and gradient vs fp64 autograd through the oracle
    d = synth.make_batch(12, 64, seed=21, outlier_frac=0.0)
    gen = torch.Generator().manual_seed(3)
    E0 = T(d["E_gt"])
    layers = [(E0 + s * torch.randn(E0.shape, generator=gen)) for s in (0.08, 0.04, 0.02)]
    ours = [e.clone().cuda().requires_grad_(True) for e in layers]
    r = get_Rt_loss(ours, T(d["Ks"]), None, None, T(d["delta_Rtijs_4_4"]), T(d["q_cam"]).cuda(), T(d["t_cam"]).cuda())
    loss = pose_loss_from_Rt_loss(r, clamp_q=0.1, clamp_t=0.5, balance_q=1.0, balance_t=0.1)
    loss.backward()
    ref_in = [e.clone().double().requires_grad_(True) for e in layers]
    lo, q_all, t_all, ra, ta = O.pose_loss(ref_in, T(d["q_cam"]).double(), T(d["t_cam"]).double(),
                                           T(d["delta_Rtijs_4_4"]).double(), 0.1, 0.5, 1.0, 0.1)
    lo.backward()
    assert abs(float(loss) - float(lo)) < 1e-5
    np.testing.assert_allclose(torch.stack(r["q_l2_error_layers_list"]).detach().cpu().numpy(), q_all.detach().numpy(), atol=2e-5)
    np.testing.assert_allclose(np.stack(r["R_angle_error_layers_list"]), ra.numpy(), atol=1e-3)
    np.testing.assert_allclose(np.stack(r["t_angle_error_layers_list"]), ta.numpy(), atol=1e-3)
    assert abs(r["R_angle_error_mean"] - float(ra.mean(1).mean())) < 1e-3
    for a, b in zip(ours, ref_in):
        rel = float((a.grad.cpu().double() - b.grad).norm() / b.grad.norm().clamp_min(1e-30))
        assert rel < 2e-3, rel


def test_deepf_training_loss_equals_the_unfused_composition():
    """losses.deepf_training_loss (one launch: F-loss on the virtual points + E = K^T T2^T F T1 K + q/t errors) against the
    composition the reference's trainer makes of get_all_loss_DeepF / get_Rt_loss (torch glue + losses.get_Rt_loss):
    same loss, same gradient w.r.t. every layer's F."""
    import torch
    from fepe_b200 import losses, ops, synth
    from oracle import fepe_oracle as O
    d = synth.make_batch(6, 300, seed=77)
    Tn = lambda k: torch.from_numpy(d[k]).cuda()
    aff = ops.hw_affine(d["image_size"])
    torch.manual_seed(1)
    ax, bx, ay, by = aff
    Tinv = torch.linalg.inv(torch.tensor([[ax, 0, bx], [0, ay, by], [0, 0, 1]], device="cuda", dtype=torch.float64))
    base = Tinv.T @ Tn("F_gt").double() @ Tinv                  # the pixel-space F_gt in DeepFNet's (H, W)-normalised frame
    base = (base / base.flatten(1).norm(dim=1).view(-1, 1, 1)).float()
    # small perturbations: most virtual points stay below the clamp, so the F-loss has a gradient
    layers = [(base + 0.002 * (l + 1) * torch.randn_like(base)).requires_grad_(True) for l in range(3)]
    Ks, Rt, q, t = Tn("Ks"), Tn("delta_Rtijs_4_4"), Tn("q_cam"), Tn("t_cam")
    v1, v2 = Tn("pts1_virt"), Tn("pts2_virt")
    loss_f, parts = losses.deepf_training_loss(layers, Ks, v1, v2, Rt, q, t, aff, clamp_at=0.02)
    g_f = torch.autograd.grad(loss_f, layers)
    # unfused: T1 = T2 = the image affine as a matrix
    T1 = torch.tensor([[ax, 0, bx], [0, ay, by], [0, 0, 1]], device="cuda").expand(6, 3, 3)
    p1 = (T1 @ v1.transpose(1, 2)).transpose(1, 2)
    p2 = (T1 @ v2.transpose(1, 2)).transpose(1, 2)
    loss_F = sum(O.epi_residual(p1, p2, Fo, 0.02).mean() for Fo in layers) / len(layers)
    TK = T1 @ Ks
    E_layers = [TK.transpose(1, 2) @ Fo @ TK for Fo in layers]
    rt = losses.get_Rt_loss(E_layers, None, None, None, Rt, q, t)
    loss_u = loss_F + losses.pose_loss_from_Rt_loss(rt)
    g_u = torch.autograd.grad(loss_u, layers)
    assert abs(float(loss_f) - float(loss_u)) < 2e-5 * abs(float(loss_u)) + 1e-7
    for a, b in zip(g_f, g_u):
        assert float((a - b).norm() / b.norm()) < 2e-3
    assert parts["q_l2"].shape == (3, 6) and parts["R_angle"].shape == (3, 6)

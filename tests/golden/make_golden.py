"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only at
/root/reference) on seeded synthetic inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The GPU box has no /root/reference, so only the committed .npz files travel; the tests
never import the reference.  Import stubs cover packages the reference imports at module
scope but never uses on this path (matplotlib, pebble, superpoint.*) -- SURVEY.md 8(c).
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "pytorch-deepfepe_b200"))
REF = "/root/reference"
sys.path[:0] = [REF, os.path.join(REF, "deepFEPE")]


def install_stubs():
    """Import stubs for packages the reference imports and never uses on this path (shared with oracle/ref_env.py)."""
    sys.path.insert(0, ROOT)
    from oracle import ref_env
    ref_env.install_stubs()


def main():
    install_stubs()
    from fepe_b200 import synth
    with contextlib.redirect_stdout(io.StringIO()):
        from deepFEPE.models.DeepFNet import Fit, NormalizeAndExpand_HW, DeepFNet
        from deepFEPE.models.ErrorEstimators import ErrorEstimator
        from deepFEPE.dsac_tools import utils_F, utils_geo
        import train_good_utils as tgu

    torch.set_num_threads(4)
    T = torch.from_numpy
    out = {}

    # ---- Fit / norm_HW / epi residual / F-loss on three weight modes ---------------------
    fit = Fit(is_cuda=False, if_cpu_svd=False)
    cases = [("uniform", 4, 200, 0), ("softmax", 4, 200, 1), ("peaked", 4, 200, 2),
             ("inlier", 4, 200, 3), ("softmax", 2, 1000, 4), ("inlier", 2, 1000, 5),
             ("softmax", 3, 37, 6)]  # ragged (N % 4 != 0)
    for i, (mode, B, N, seed) in enumerate(cases):
        d = synth.make_batch(B, N, seed, weight_mode=mode)
        nhw = NormalizeAndExpand_HW(d["image_size"], is_cuda=False)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            p1, p2, T1, T2 = nhw(T(d["matches_xy_ori"]))
            p1, p2 = p1.permute(0, 2, 1).contiguous(), p2.permute(0, 2, 1).contiguous()
            Fo, res = fit(p1, p2, T(d["weights"]))
            epi = utils_F.compute_epi_residual(p1, p2, Fo)
            v1, v2, Ks = T(d["pts1_virt"]), T(d["pts2_virt"]), T(d["Ks"])
            pe1 = (T1 @ v1.permute(0, 2, 1)).permute(0, 2, 1)
            pe2 = (T2 @ v2.permute(0, 2, 1)).permute(0, 2, 1)
            lossF = utils_F.compute_epi_residual(pe1, pe2, Fo, 0.02)
            E = Ks.transpose(1, 2) @ T2.permute(0, 2, 1) @ Fo @ T1 @ Ks
        out[f"fit{i}_meta"] = np.array([B, N, seed])
        out[f"fit{i}_mode"] = np.array(mode)
        out[f"fit{i}_pts1"], out[f"fit{i}_pts2"] = p1.numpy(), p2.numpy()
        out[f"fit{i}_T1"] = T1.contiguous().numpy()
        out[f"fit{i}_F"], out[f"fit{i}_res"], out[f"fit{i}_epi"] = Fo.numpy(), res.numpy(), epi.numpy()
        out[f"fit{i}_lossF"], out[f"fit{i}_E"] = lossF.numpy(), E.numpy()

    # ---- Fit backward (autograd through torch.svd) ---------------------------------------
    d = synth.make_batch(3, 120, 11, weight_mode="inlier")
    nhw = NormalizeAndExpand_HW(d["image_size"], is_cuda=False)
    with contextlib.redirect_stdout(io.StringIO()):
        p1, p2, T1, T2 = nhw(T(d["matches_xy_ori"]))
        p1 = p1.permute(0, 2, 1).contiguous().double()
        p2 = p2.permute(0, 2, 1).contiguous().double()
        w = T(d["weights"]).double().requires_grad_(True)
        fit64 = Fit(is_cuda=False, if_cpu_svd=False).double()
        fit64.ones_b, fit64.T_b, fit64.mask = fit64.ones_b.double(), fit64.T_b.double(), fit64.mask.double()
        Fo, res = fit64(p1, p2, w)
        epi = utils_F.compute_epi_residual(p1, p2, Fo)
        g = torch.Generator().manual_seed(5)
        gF = torch.randn(Fo.shape, generator=g, dtype=torch.float64)
        gr = torch.randn(res.shape, generator=g, dtype=torch.float64)
        ge = torch.randn(epi.shape, generator=g, dtype=torch.float64)
        # the sign of f is LAPACK's; make the scalar sign-invariant the way tests do: through F*sign
        sgn = torch.sign(Fo.detach()[:, 2, 2]).view(-1, 1, 1)
        (((Fo * sgn) * gF).sum() + ((res * sgn.view(-1, 1)) * gr).sum() + (epi * ge).sum()).backward()
    out["bwd_pts1"], out["bwd_pts2"], out["bwd_w"] = p1.numpy(), p2.numpy(), w.detach().numpy()
    out["bwd_gF"], out["bwd_gr"], out["bwd_ge"] = gF.numpy(), gr.numpy(), ge.numpy()
    out["bwd_F"], out["bwd_res"], out["bwd_epi"] = Fo.detach().numpy(), res.detach().numpy(), epi.detach().numpy()
    out["bwd_gw"] = w.grad.numpy()

    # ---- pose: _get_M2s, _R_to_q, get_Rt_loss --------------------------------------------
    d = synth.make_batch(6, 300, 21, weight_mode="inlier", outlier_frac=0.1)
    nhw = NormalizeAndExpand_HW(d["image_size"], is_cuda=False)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        p1, p2, T1, T2 = nhw(T(d["matches_xy_ori"]))
        p1, p2 = p1.permute(0, 2, 1).contiguous(), p2.permute(0, 2, 1).contiguous()
        Fo, _ = fit(p1, p2, T(d["weights"]))
        Ks = T(d["Ks"])
        E = Ks.transpose(1, 2) @ T2.permute(0, 2, 1) @ Fo @ T1 @ Ks
        R1s, R2s, ts, q1s, q2s = [], [], [], [], []
        for Ec in E.transpose(1, 2):
            Rs, tt, _ = utils_F._get_M2s(Ec)
            R1s.append(Rs[0]), R2s.append(Rs[1]), ts.append(tt[0])
            q1s.append(utils_geo._R_to_q(Rs[0])), q2s.append(utils_geo._R_to_q(Rs[1]))
        res = tgu.get_Rt_loss([E, E * 1.0], Ks, T(d["matches_xy_ori"][:, :, :2]),
                              T(d["matches_xy_ori"][:, :, 2:]), T(d["delta_Rtijs_4_4"]),
                              T(d["q_cam"]), T(d["t_cam"]), device="cpu")
    out["pose_E"], out["pose_Rt"] = E.numpy(), d["delta_Rtijs_4_4"]
    out["pose_qcam"], out["pose_tcam"] = d["q_cam"], d["t_cam"]
    out["pose_R1"], out["pose_R2"] = torch.stack(R1s).numpy(), torch.stack(R2s).numpy()
    out["pose_t"] = torch.stack(ts).numpy()
    out["pose_q1"], out["pose_q2"] = torch.stack(q1s).numpy(), torch.stack(q2s).numpy()
    out["pose_q_l2"] = torch.stack(res["q_l2_error_layers_list"]).numpy()
    out["pose_t_l2"] = torch.stack(res["t_l2_error_layers_list"]).numpy()
    out["pose_R_ang"] = np.stack(res["R_angle_error_layers_list"])
    out["pose_t_ang"] = np.stack(res["t_angle_error_layers_list"])

    # quaternion branches: rotations about each axis by ~pi exercise all four
    Rq = synth.rodrigues(np.array([[3.0, 0.1, 0.0], [0.1, 3.0, 0.0], [0.0, 0.1, 3.0], [0.2, -0.1, 0.3],
                                   [2.2, 2.2, 0.1], [0.0, 2.2, 2.2]]))
    out["quat_R"] = Rq.astype(np.float32)
    out["quat_q"] = np.stack([utils_geo._R_to_q(T(r.astype(np.float32))).numpy() for r in Rq])

    # ---- ErrorEstimator + full DeepFNet forward (config 1 plumbing) -----------------------
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        ee = ErrorEstimator(4)
    x = torch.rand(2, 4, 64, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        y = ee(x)
    out["ee_x"], out["ee_y"] = x.numpy(), y.numpy()
    out["ee_keys"] = np.array(list(ee.state_dict().keys()))

    cuda_backup = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self      # DeepFNet.__init__ calls .cuda() (:356)
    try:
        torch.manual_seed(77)
        with contextlib.redirect_stdout(io.StringIO()):
            net = DeepFNet(depth=5, image_size=[376, 1241, 3], if_quality=False, is_cuda=False,
                           if_cpu_svd=False)
        d = synth.make_batch(1, 100, 31, planar=True, outlier_frac=0.0)
        d2 = synth.make_batch(2, 160, 32, weight_mode="softmax")
        for tag, dd in (("c1", d), ("c1b", d2)):
            batch = {"matches_xy_ori": T(dd["matches_xy_ori"]),
                     "matches_good_unique_nums": T(dd["matches_good_unique_nums"]),
                     "t_scene_scale": torch.ones(dd["matches_xy_ori"].shape[0], 1, 1)}
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                o = net(batch)
            out[f"{tag}_matches"] = dd["matches_xy_ori"]
            out[f"{tag}_F_layers"] = torch.stack(o["out_layers"]).numpy()
            out[f"{tag}_res_layers"] = torch.stack(o["residual_layers"]).numpy()
            out[f"{tag}_epi_layers"] = torch.stack(o["epi_res_layers"]).numpy()
            out[f"{tag}_w_layers"] = torch.stack(o["weights_layers"]).numpy()
            out[f"{tag}_logits"] = o["logits"].numpy()
        out["c1_state_keys"] = np.array(list(net.state_dict().keys()))
    finally:
        torch.Tensor.cuda = cuda_backup

    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

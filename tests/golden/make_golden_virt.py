"""Generate tests/golden/gt_virt_ref.npz: the UNMODIFIED reference functions the dataset calls per sample for the
ground-truth side of a pair -- utils_F.E_F_from_Rt_np (dsac_tools/utils_F.py:835-846), utils_misc.get_virt_x1x2_grid /
get_virt_x1x2_np (dsac_tools/utils_misc.py:163-199, cv2.correctMatches inside; opencv 4.13 in this image) and
utils_geo.R_to_q_np (dsac_tools/utils_geo.py:88-117) -- in the order of deepFEPE/datasets/kitti_odo_corr.py:290-302 and
:526-566, on seeded synthetic scene motions with the float32 dtypes the dataset hands them.  Also stores raw
cv2.correctMatches outputs (NaNs kept) for random rank-2 F and points, which pin the polynomial-solver restatement.
Pins oracle/virt_points_oracle.py and, through it, fepe_gt_virt.
Run once:  python tests/golden/make_golden_virt.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (path setup + import stubs)


def main():
    import cv2
    MG.install_stubs()
    from fepe_b200 import synth
    with contextlib.redirect_stdout(io.StringIO()):
        from deepFEPE.dsac_tools import utils_F, utils_geo, utils_misc
    image_size = synth.KITTI_IMAGE_SIZE
    g1, g2 = utils_misc.get_virt_x1x2_grid(image_size)
    keys = ("K", "Rt", "E", "F", "pts1_virt_normalized", "pts2_virt_normalized", "pts1_virt", "pts2_virt",
            "q_cam", "t_cam", "q_scene", "t_scene")
    rec = {k: [] for k in keys}
    for seed in range(7):
        # seeds 0-2: forward-dominant KITTI-like motion (epipole inside the image); 3-5: sideways motion;
        # 6: a batch in which OpenCV returns NaN for a grid point next to the epipole (the reference stores 0)
        d = synth.make_batch(4, 8, seed=300 + seed) if seed < 6 else synth.make_batch(8, 16, seed=3)
        for b in range(4):
            K = d["Ks"][b]                                   # float32, like cam.npy after astype (kitti_odo_corr.py:115)
            Rt = d["delta_Rtijs_4_4"][b].copy()              # float32 scene motion
            if 3 <= seed < 6:
                rng = np.random.default_rng(seed * 10 + b)
                t = rng.normal(size=3).astype(np.float32)
                Rt[:3, 3] = t / np.linalg.norm(t)
            E, F = utils_F.E_F_from_Rt_np(Rt[:3, :3], Rt[:3, 3:4], K)
            n1, n2, p1, p2 = utils_misc.get_virt_x1x2_np(image_size, F, K, g1, g2)
            Rt_cam = np.linalg.inv(Rt)
            vals = (K, Rt, E, F, n1, n2, p1, p2, utils_geo.R_to_q_np(Rt_cam[:3, :3]), Rt_cam[:3, 3:4],
                    utils_geo.R_to_q_np(Rt[:3, :3]), Rt[:3, 3:4])
            for k, v in zip(keys, vals):
                rec[k].append(np.asarray(v))
    out = {k: np.stack(v) for k, v in rec.items()}
    out["grid1"], out["grid2"] = g1, g2

    rng = np.random.default_rng(11)
    Fs, P1, P2, C1, C2 = [], [], [], [], []
    for k in range(12):
        A = rng.normal(size=(3, 3))
        U, S, Vt = np.linalg.svd(A)
        S[2] = 0.0
        sc = [1.0, 1e-2, 1e-4][k % 3]                       # from normalised to pixel-like scaling of F
        r = np.sqrt(sc)
        F = (U @ np.diag(S) @ Vt) * np.array([[sc, sc, r], [sc, sc, r], [r, r, 1.0]])
        p1 = (rng.uniform(-1, 1, size=(25, 2)) / r).astype(np.float32)
        p2 = (rng.uniform(-1, 1, size=(25, 2)) / r).astype(np.float32)
        c1, c2 = cv2.correctMatches(F, p1[None], p2[None])
        Fs.append(F), P1.append(p1), P2.append(p2), C1.append(c1[0]), C2.append(c2[0])
    out.update(cv_F=np.stack(Fs), cv_p1=np.stack(P1), cv_p2=np.stack(P2), cv_c1=np.stack(C1), cv_c2=np.stack(C2))
    out["cv2_version"] = np.array(cv2.__version__)
    path = os.path.join(HERE, "gt_virt_ref.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {out['K'].shape[0]} samples + {len(Fs)} raw cv2 cases, {os.path.getsize(path) / 1024:.0f} KiB;"
          f" NaN points in raw cases: {int(np.isnan(out['cv_c1'][..., 0]).sum())},"
          f" zeroed virtual points: {int((out['pts1_virt'][..., :2] == 0).all(-1).sum())}")
    for k in keys:
        print(f"  {k:22s} {out[k].dtype} {out[k].shape}")


if __name__ == "__main__":
    main()

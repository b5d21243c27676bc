"""Generate tests/golden/recover_pose_ref.npz: the UNMODIFIED reference function
deepFEPE/dsac_tools/utils_F.py:909-954 goodCorr_eval_nondecompose (cv2.recoverPose + utils_geo error metrics), and
cv2.recoverPose itself (opencv 4.13 in this image), on seeded synthetic scenes -- exact and perturbed essential
matrices, 30 % outliers.  Pins oracle/recover_pose_oracle.py and, through it, fepe_recover_pose.
Run once:  python tests/golden/make_golden_recover.py
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
        from deepFEPE.dsac_tools import utils_F
    rng = np.random.default_rng(7)
    rec = {k: [] for k in ("E", "K", "matches", "Rt", "M", "err", "good", "mask")}
    for seed in range(4):
        d = synth.make_batch(4, 300, seed=100 + seed)
        for b in range(4):
            K = d["Ks"][b].astype(np.float64)
            m = d["matches_xy_ori"][b]
            x1, x2 = m[:, :2].astype(np.float64), m[:, 2:].astype(np.float64)
            Rt = d["delta_Rtijs_4_4"][b].astype(np.float64)
            inv = np.linalg.inv(Rt)[:3]
            for noise in (0.0, 0.03):
                E = (d["E_gt"][b] + noise * rng.normal(size=(3, 3))).astype(np.float32)   # what the model hands over
                with contextlib.redirect_stdout(io.StringIO()):
                    M, err = utils_F.goodCorr_eval_nondecompose(x1, x2, E.astype(np.float64), inv, K, None)
                good, _, _, mask = cv2.recoverPose(E.astype(np.float64), x1, x2, focal=float(K[0, 0]),
                                                   pp=(float(K[0, 2]), float(K[1, 2])))
                rec["E"].append(E), rec["K"].append(d["Ks"][b]), rec["matches"].append(m), rec["Rt"].append(d["delta_Rtijs_4_4"][b])
                rec["M"].append(M), rec["err"].append(np.array(err)), rec["good"].append(good)
                rec["mask"].append(mask.reshape(-1))
    out = {k: np.stack(v) for k, v in rec.items()}
    out["cv2_version"] = np.array(cv2.__version__)
    path = os.path.join(HERE, "recover_pose_ref.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {out['E'].shape[0]} cases, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

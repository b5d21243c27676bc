"""Generate tests/golden/dump_ref.npz: a small synthetic correspondence dump in the reference's .npy wire format
(deepFEPE_data/dump_tools/kitti_seq_loader.py:351-357, :614-632) together with what the UNMODIFIED reference dataset
class deepFEPE/datasets/kitti_odo_corr.py::KittiCorrOdo returns for every sample of it (numpy global RNG seeded), with
the config values of deepFEPE/configs/kitti_corr_baseline.yaml (resize [376,1240], good_num 1000, with_quality).
Pins fepe_b200.dumps.KittiCorrDump and, for the ground-truth keys, fepe_gt_virt.
The files are written here with plain np.save calls, not with the product's writer.
Run once:  python tests/golden/make_golden_dump.py
"""
import collections
import collections.abc
import contextlib
import io
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (path setup + import stubs)

SCENE = "04_02"
CONFIG = {"sequence_length": 2, "delta_ij": 1, "batch_size": 4, "good_num": 1000,
          "read_what": {"with_X": False, "with_pose": True, "with_sift": True, "with_sift_des": False, "with_SP": False,
                        "with_quality": True, "with_qt": True, "with_imgs": False},
          "image": {"size": [376, 1241, 3]}, "preprocessing": {"resize": [376, 1240]}, "read_params": {"use_h5": False}}
SEED = 123


class _Path(str):               # the few members of path.Path the dataset uses
    def __truediv__(self, o):
        return _Path(os.path.join(str(self), str(o)))

    def __add__(self, o):
        return _Path(str.__add__(self, o))

    @property
    def name(self):
        return os.path.basename(str(self))

    def isfile(self):
        return os.path.isfile(str(self))


def main():
    MG.install_stubs()
    MG._stub("path", Path=_Path)
    MG._stub("imageio", imread=None)
    MG._stub("skimage")
    MG._stub("skimage.transform", resize=None)
    MG._stub("pykitti")
    MG._stub("coloredlogs", install=lambda *a, **k: None)
    MG._stub("termcolor", colored=lambda t, *a, **k: t, cprint=print)
    collections.Mapping = collections.abc.Mapping      # python < 3.10 alias the reference uses (utils/tools.py:18)
    from fepe_b200 import synth

    rng = np.random.default_rng(0)
    nfr = 4
    files = {"cam": synth.KITTI_K.astype(np.float32)}
    Rt_cam2 = np.eye(4, dtype=np.float32)
    Rt_cam2[:3, 3] = [0.06, -0.01, 0.002]
    files["Rt_cam2_gt"] = Rt_cam2
    poses, T = [], np.eye(4)
    for _ in range(nfr):
        poses.append(T[:3].copy())
        d = np.eye(4)
        d[:3, :3] = synth.rodrigues(rng.normal(0, 0.02, size=(1, 3)))[0]
        d[:3, 3] = [0.02 * rng.normal(), 0.01 * rng.normal(), 0.9 + 0.1 * rng.random()]
        T = T @ d
    files["poses"] = np.stack(poses).astype(np.float32).reshape(nfr, -1)
    sizes = [(2100, 1200), (400, 300), (900, 900)]     # crop both / pad both / pad all
    for i, (n_all, n_good) in enumerate(sizes):
        m = np.c_[rng.uniform(0, 1241, n_all), rng.uniform(0, 376, n_all), rng.uniform(0, 1241, n_all),
                  rng.uniform(0, 376, n_all), rng.uniform(50, 300, n_all), rng.uniform(0.3, 1, n_all)].astype(np.float32)
        m[5] = m[4]                                     # a duplicated row: matches_all_unique_nums < M
        files[f"ij_match_quality_{i}-{i + 1}_all"] = m
        files[f"ij_match_quality_{i}-{i + 1}_good"] = m[:n_good].copy()

    root = tempfile.mkdtemp(prefix="fepe_dump_")
    try:
        os.makedirs(os.path.join(root, SCENE))
        for k, v in files.items():
            np.save(os.path.join(root, SCENE, k + ".npy"), v)
        for k in range(nfr):                            # the dataset only checks that these exist (:152-160)
            open(os.path.join(root, SCENE, f"{k:06d}.jpg"), "wb").close()
            np.save(os.path.join(root, SCENE, f"sift_{k:06d}.npy"), np.zeros((1, 130), np.float32))
        with open(os.path.join(root, "train.txt"), "w") as f:
            for i in range(len(sizes)):
                f.write(f"{SCENE} {i:06d}\n")
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            from deepFEPE.datasets.kitti_odo_corr import KittiCorrOdo
            cfg = {"data": dict(CONFIG, dump_root=root),
                   "model": {"if_img_feat": False, "if_SP": False, "if_lidar_corres": False}}
            ds = KittiCorrOdo(task="train", seed=0, **cfg)
            np.random.seed(SEED)
            samples = [ds[i] for i in range(len(ds))]
    finally:
        shutil.rmtree(root)

    # the _good files are the first n_good rows of the _all files: store their sizes only
    out = {"file_" + k: v for k, v in files.items() if not k.endswith("_good")}
    out["good_sizes"] = np.array([g for _, g in sizes])
    keys = ("K_ori", "K", "K_inv", "E", "F", "Rt_cam2_gt", "matches_all", "matches_good", "quality_good",
            "pts1_virt_normalized", "pts2_virt_normalized", "pts1_virt", "pts2_virt", "q_cam", "t_cam", "q_scene", "t_scene")
    for k in keys:
        out["ref_" + k] = np.stack([s[k] for s in samples])
    assert all(s["quality_all"] is s["quality_good"] or np.array_equal(s["quality_all"], s["quality_good"]) for s in samples)
    out["ref_matches_good_unique_nums"] = np.array([s["matches_good_unique_nums"] for s in samples])
    out["ref_matches_all_unique_nums"] = np.array([s["matches_all_unique_nums"] for s in samples])
    out["ref_relative_scene_pose"] = np.stack([s["relative_scene_poses"][1] for s in samples])
    out["ref_cam_poses"] = np.stack([np.stack(s["cam_poses"]) for s in samples])
    out["ref_frame_ids"] = np.array([s["frame_ids"] for s in samples])
    out["seed"] = np.array(SEED)
    path = os.path.join(HERE, "dump_ref.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(samples)} samples, {os.path.getsize(path) / 1024:.0f} KiB")
    for k in sorted(out):
        if k.startswith("ref_"):
            print(f"  {k:32s} {out[k].dtype} {out[k].shape}")


if __name__ == "__main__":
    main()

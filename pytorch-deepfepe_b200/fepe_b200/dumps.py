"""Reader (and writer) of the reference's ``.npy`` correspondence dump -- SURVEY.md 8f rank 4.

Wire format, written by deepFEPE_data/dump_tools/kitti_seq_loader.py (:351-357 per scene, :614-632 per frame pair) and
read by deepFEPE/datasets/kitti_odo_corr.py (KittiCorrOdo, :112-145, :340-349, :451-509):

    <dump_root>/{train,val,test}.txt            one line per first frame: "<scene> <%06d frame>"   (:64-67)
    <dump_root>/<scene>/cam.npy                 float32 [3,3]   intrinsics of the dumped (unresized) images
    <dump_root>/<scene>/poses.npy               float32 [n,12]  absolute pose [R|t] of every frame, row-major 3x4
    <dump_root>/<scene>/Rt_cam2_gt.npy          float32 [4,4]   cam0 -> cam2 extrinsics (scenes ending in "02" conjugate
                                                                the relative motion with it, :199-206)
    <dump_root>/<scene>/ij_match_quality_{i}-{j}_all.npy   float32 [M,6]  x1,y1,x2,y2, descriptor distance, ratio
    <dump_root>/<scene>/ij_match_quality_{i}-{j}_good.npy  float32 [G,6]  the ratio-test survivors, same columns

`KittiCorrDump.sample()` reproduces the keys KittiCorrOdo.__getitem__ derives from those files bit for bit (same numpy
calls on the same dtypes, same order of draws from the numpy global RNG for the crop / pad choice, :452-474 ->
dsac_tools/utils_misc.py:139-161).  `KittiCorrDump.batch()` collates samples, uploads them once and builds everything
the reference computed per sample on the host AFTER reading -- E, F, virtual correspondences, q / t
(:290-302, :526-566) -- on the device with fepe_gt_virt (fepe_b200.gt).  `data_batch()` then shapes the dict
DeepFNet.forward and the loss read (deepFEPE/Train_model_pipeline.py:330-446).  Images, SIFT descriptors and lidar
points of the dump are not read: the path never uses them (with_imgs / with_sift_des / with_X are false in
deepFEPE/configs/kitti_corr_baseline.yaml:14-22).
"""
from __future__ import annotations

import os
from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from . import gt as _gt

ALL_NUM = 2000           # kitti_odo_corr.py:455: matches_all is always cropped / padded to 2000


def crop_or_pad_choice(in_num_points: int, out_num_points: int, shuffle: bool = False, rng=np.random) -> np.ndarray:
    """dsac_tools/utils_misc.py:139-161: indices that crop or (with replacement) pad a set to a fixed size.
    `rng` is the numpy global RNG module by default, like the reference, so np.random.seed() reproduces its draws."""
    choice = rng.permutation(in_num_points) if shuffle else np.arange(in_num_points)
    assert out_num_points > 0, "out_num_points = %d must be positive int!" % out_num_points
    if in_num_points >= out_num_points:
        return choice[:out_num_points]
    pad = rng.choice(choice, out_num_points - in_num_points, replace=True)
    return np.concatenate([choice, pad])


def _rt_pad(Rt: np.ndarray) -> np.ndarray:
    """dsac_tools/utils_misc.py:96-99."""
    return np.vstack((Rt, np.array([[0.0, 0.0, 0.0, 1.0]], dtype=Rt.dtype)))


def write_dump_scene(dump_root: str, scene: str, K: np.ndarray, poses: np.ndarray, Rt_cam2_gt: np.ndarray,
                     pairs: dict, split: str = "train") -> None:
    """Write one scene in the reference's layout (kitti_seq_loader.py:351-357, :614-632).  `poses` [n,3,4];
    `pairs` maps (i, j) -> (match_quality_all [M,6], match_quality_good [G,6]).  Appends the first frames to
    <split>.txt.  Used to put synthetic scenes on disk for tests and benchmarks; real dumps come from the reference's
    own dump tool."""
    d = os.path.join(dump_root, scene)
    os.makedirs(d, exist_ok=True)
    np.save(os.path.join(d, "cam.npy"), np.asarray(K).astype(np.float32))
    np.save(os.path.join(d, "Rt_cam2_gt.npy"), np.asarray(Rt_cam2_gt).astype(np.float32))
    poses = np.asarray(poses)
    np.save(os.path.join(d, "poses.npy"), poses.reshape(poses.shape[0], -1).astype(np.float32))
    with open(os.path.join(dump_root, f"{split}.txt"), "a") as f:
        for (i, j), (m_all, m_good) in sorted(pairs.items()):
            np.save(os.path.join(d, f"ij_match_quality_{i}-{j}_all.npy"), np.asarray(m_all))
            np.save(os.path.join(d, f"ij_match_quality_{i}-{j}_good.npy"), np.asarray(m_good))
            f.write(f"{scene} {i:06d}\n")


class KittiCorrDump:
    """The correspondence / pose part of the reference's KittiCorrOdo dataset (sequence_length 2, .npy dumps).

    dump_root, task        : as the reference's config["data"]["dump_root"] and the split file <task>.txt
    delta_ij               : frame distance of a pair (config data.delta_ij)
    good_num               : crop / pad size of matches_good (config data.good_num, 1000)
    image_size             : size the dump was made at (config data.image.size, [376,1241,3])
    resize                 : config data.preprocessing.resize ([376,1240]) or None; scales K, the matches and the grid
    with_quality           : config data.read_what.with_quality
    """

    def __init__(self, dump_root: str, task: str = "train", delta_ij: int = 1, good_num: int = 1000,
                 image_size: Sequence[int] = (376, 1241, 3), resize: Optional[Sequence[int]] = None,
                 with_quality: bool = True):
        assert task in ("train", "val", "test")
        self.root, self.task, self.delta_ij, self.good_num = str(dump_root), task, int(delta_ij), int(good_num)
        self.image_size = list(image_size)
        self.sizerHW = list(resize) if resize else list(image_size)                    # kitti_odo_corr.py:85-91
        self.with_quality = with_quality
        # :272: without images the zoom follows from the two sizes
        self.zoom_xy = (self.sizerHW[1] / self.image_size[1], self.sizerHW[0] / self.image_size[0])
        frames = []
        with open(os.path.join(self.root, f"{task}.txt")) as f:
            for line in f:                                                              # :64-67: "<scene> <%06d>\n"
                if line.strip():
                    frames.append((line[:-8], line[-7:-1]))
        cam_ids = {s[-2:] for s, _ in frames}
        if len(cam_ids) != 1:
            raise ValueError(f"{task}.txt mixes cameras {sorted(cam_ids)} (kitti_odo_corr.py:70-71)")
        self.cam_id = cam_ids.pop()
        self._K, self._poses, self._Rt_cam2 = {}, {}, {}
        self.samples = []
        for scene, frame_id in frames:                                                  # crawl_folders, :100-221
            d = os.path.join(self.root, scene)
            if scene not in self._K:
                self._K[scene] = np.load(os.path.join(d, "cam.npy")).astype(np.float32).reshape((3, 3))
                self._poses[scene] = np.load(os.path.join(d, "poses.npy")).astype(np.float32).reshape(-1, 3, 4)
                self._Rt_cam2[scene] = np.load(os.path.join(d, "Rt_cam2_gt.npy"))
            i = int(frame_id)
            j = i + self.delta_ij
            if not os.path.isfile(os.path.join(d, f"ij_match_quality_{i}-{j}_good.npy")):
                continue                                                                # :140-145: skipped with a warning
            poses, Rt_cam2 = self._poses[scene], self._Rt_cam2[scene]
            rel = np.linalg.inv(_rt_pad(poses[j])) @ _rt_pad(poses[i])                  # :196-198
            if self.cam_id == "02":
                rel = Rt_cam2 @ rel @ np.linalg.inv(Rt_cam2)                            # :199-204
            self.samples.append({
                "scene": scene, "scene_name": task + os.path.basename(scene), "ids": [i, j],
                "frame_ids": ["%06d" % i, "%06d" % j], "K_ori": self._K[scene], "Rt_cam2_gt": Rt_cam2,
                "cam_poses": [np.linalg.inv(_rt_pad(poses[i])), np.linalg.inv(_rt_pad(poses[j]))],      # :190-192
                "relative_scene_poses": [np.hstack((np.eye(3, dtype=np.float32), np.zeros((3, 1), dtype=np.float32))),
                                         rel],                                                          # :193-209
            })

    def __len__(self) -> int:
        return len(self.samples)

    def sample(self, index: int, rng=np.random) -> dict:
        """The file-derived keys of KittiCorrOdo.__getitem__ (numpy, host): K_ori, K, K_inv, matches_all [2000,4],
        matches_good [good_num,4], matches_good_unique_nums, matches_all_unique_nums, quality_good / quality_all
        [good_num,2], relative_scene_poses, cam_poses, Rt_cam2_gt, frame_ids, scene_name.  Bit-identical to the
        reference given the same numpy RNG state."""
        s = self.samples[index]
        zx, zy = self.zoom_xy
        out = {k: s[k] for k in ("K_ori", "scene_name", "frame_ids", "Rt_cam2_gt", "relative_scene_poses", "cam_poses")}
        P = np.concatenate((s["K_ori"], [[0], [0], [0]]), axis=1).astype(np.float32)   # add_scaled_K, :276-288
        P[0] *= zx
        P[1] *= zy
        K = P[:, :3]
        out["K"], out["K_inv"] = K, np.linalg.inv(K)
        base = os.path.join(self.root, s["scene"], "ij_match_quality_{}-{}".format(*s["ids"]))
        mq_all = np.load(base + "_all.npy").astype(np.float32)                          # :440-446
        mq_good = np.load(base + "_good.npy").astype(np.float32)

        def scaled(m):                                                                  # scale_points(.., loop_length=4), :304-311
            m = m[:, :4]                                                                # a view: the reference scales in place
            for c in range(4):
                m[:, c] = m[:, c] * (zx, zy)[c % 2]
            return m

        m_all = scaled(mq_all)
        choice_all = crop_or_pad_choice(m_all.shape[0], ALL_NUM, shuffle=True, rng=rng)               # :454-456
        m_good = scaled(mq_good)
        choice_good = crop_or_pad_choice(m_good.shape[0], self.good_num, shuffle=True, rng=rng)        # :462-466
        out.update({"matches_all": m_all[choice_all], "matches_good": m_good[choice_good],
                    "matches_good_unique_nums": min(m_good.shape[0], self.good_num),
                    "matches_all_unique_nums": np.unique(m_all, axis=0).shape[0]})                      # :468-478
        if self.with_quality:                                                                           # :482-509
            q = mq_good[:, 4:][choice_good]
            q[:, 0] = q[:, 0] / 300.0
            out.update({"quality_good": q, "quality_all": q})
        out["get_flags"] = {"have_matches": True}
        return out

    def batch(self, indices: Iterable[int], device="cuda", rng=np.random) -> dict:
        """Collated samples on `device` (what the DataLoader's default collate + the .cuda() calls of
        Train_model_pipeline.py:330-446 give), plus the ground-truth keys built ON the device:
        E, F, pts{1,2}_virt, pts{1,2}_virt_normalized, q_cam, t_cam, q_scene, t_scene (fepe_b200.gt.gt_sample_batch)."""
        ss = [self.sample(i, rng) for i in indices]
        dev = torch.device(device)
        up = lambda key: torch.from_numpy(np.stack([s[key] for s in ss])).to(dev, non_blocking=True)
        out = {k: up(k) for k in ("K_ori", "K", "K_inv", "matches_all", "matches_good", "Rt_cam2_gt")}
        if self.with_quality:
            out["quality_good"] = up("quality_good")
            out["quality_all"] = out["quality_good"]
        out["matches_good_unique_nums"] = torch.tensor([s["matches_good_unique_nums"] for s in ss])
        out["matches_all_unique_nums"] = torch.tensor([s["matches_all_unique_nums"] for s in ss])
        out["relative_scene_poses"] = [
            torch.from_numpy(np.stack([s["relative_scene_poses"][0] for s in ss])).to(dev),
            torch.from_numpy(np.stack([s["relative_scene_poses"][1] for s in ss]).astype(np.float32)).to(dev)]
        out["cam_poses"] = [torch.from_numpy(np.stack([s["cam_poses"][k] for s in ss])).to(dev) for k in range(2)]
        out["frame_ids"] = [[s["frame_ids"][k] for s in ss] for k in range(2)]
        out["scene_name"] = [s["scene_name"] for s in ss]
        grids = _gt.get_virt_x1x2_grid(self.sizerHW, device=dev)                        # :93-96: the grid of the resized image
        out.update(_gt.gt_sample_batch(out["relative_scene_poses"][1], out["K"], self.sizerHW, grids=grids))
        return out

    @staticmethod
    def data_batch(batch: dict, if_quality: bool = True) -> dict:
        """The dict DeepFNet.forward and get_all_loss_DeepF read, from a `batch()` (Train_model_pipeline.py:340-446,
        SIFT branch: matches_use = matches_good)."""
        m = batch["matches_good"]
        Ks, K_invs = batch["K"], batch["K_inv"]
        ones = torch.ones_like(m[:, :, :1])
        nrm = lambda x: (torch.cat((x, ones), 2) @ K_invs.transpose(1, 2))[:, :, :2]    # _de_homo(K^-1 _homo(x)), third row of K^-1 is (0,0,1)
        x1n, x2n = nrm(m[:, :, :2]), nrm(m[:, :, 2:])
        ts_scene = batch["t_scene"]
        return {"matches_xy": torch.cat((x1n, x2n), 2), "matches_xy_ori": m,
                "quality": batch["quality_good"] if if_quality else None,
                "x1_normalizedK": x1n, "x2_normalizedK": x2n, "Ks": Ks, "K_invs": K_invs, "des1": None, "des2": None,
                "matches_good_unique_nums": batch["matches_good_unique_nums"],
                "t_scene_scale": torch.norm(ts_scene, p=2, dim=1, keepdim=True), "frame_ids": batch["frame_ids"],
                "E_gts": batch["E"], "F_gts": batch["F"], "pts1_virt_ori": batch["pts1_virt"],
                "pts2_virt_ori": batch["pts2_virt"], "delta_Rtijs_4_4": batch["relative_scene_poses"][1],
                "qs_cam": batch["q_cam"], "ts_cam": batch["t_cam"], "qs_scene": batch["q_scene"], "ts_scene": ts_scene}

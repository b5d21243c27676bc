"""Seeded synthetic two-view batches shaped like the reference's KITTI sample dict.

The reference ships no data and no intrinsics literal (it reads ``cam.npy``,
deepFEPE/datasets/kitti_odo_corr.py:114-116), so benchmark / parity inputs are
generated here following SURVEY.md section 8(d):

* K: KITTI-odometry shaped, fx=fy=718.856, cx=607.1928, cy=185.2157, 1241x376 image
  (image size from deepFEPE/configs/kitti_corr_baseline.yaml:24).
* scene motion ``x2 = R x1 + t`` (the reference's ``delta_Rtijs_4_4`` convention,
  deepFEPE/Train_model_pipeline.py:347), R from a small rotation vector, t unit and
  forward dominant.
* 3-D points uniform in a KITTI-like frustum, projected into both views, pixel noise,
  the first ``outlier_frac`` of the second-view points replaced by uniform clutter.
* ``q_cam / t_cam`` are taken from the INVERSE motion like the dataset does
  (deepFEPE/datasets/kitti_odo_corr.py:547-554).
* virtual points: 10x10 grid (deepFEPE/dsac_tools/utils_misc.py:163-171) moved onto the
  ground-truth epipolar geometry.  The reference uses cv2.correctMatches
  (utils_misc.py:176); here the second point is simply projected onto the epipolar line
  of the first, which also gives ``x2^T F x1 = 0`` and needs no OpenCV.

Everything is numpy float64 internally and returned as float32 (the reference casts
its samples to float32, kitti_odo_corr.py:444).
"""
from __future__ import annotations

import numpy as np

KITTI_K = np.array([[718.856, 0.0, 607.1928],
                    [0.0, 718.856, 185.2157],
                    [0.0, 0.0, 1.0]], dtype=np.float64)
KITTI_IMAGE_SIZE = (376, 1241, 3)  # H, W, C as in the reference config


def rodrigues(rvec: np.ndarray) -> np.ndarray:
    """Rotation matrices from rotation vectors, rvec [B,3] -> [B,3,3]."""
    theta = np.linalg.norm(rvec, axis=-1, keepdims=True)
    k = rvec / np.maximum(theta, 1e-300)
    K = np.zeros(rvec.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    th = theta[..., None]
    eye = np.broadcast_to(np.eye(3), K.shape)
    return eye + np.sin(th) * K + (1.0 - np.cos(th)) * (K @ K)


def skew(t: np.ndarray) -> np.ndarray:
    """[B,3] -> [B,3,3] cross-product matrices."""
    S = np.zeros(t.shape[:-1] + (3, 3))
    S[..., 0, 1], S[..., 0, 2] = -t[..., 2], t[..., 1]
    S[..., 1, 0], S[..., 1, 2] = t[..., 2], -t[..., 0]
    S[..., 2, 0], S[..., 2, 1] = -t[..., 1], t[..., 0]
    return S


def rot_to_quat(R: np.ndarray) -> np.ndarray:
    """Batched trace-method quaternion (w,x,y,z), w>=0, same branch rule as the
    reference's R_to_q_np (deepFEPE/dsac_tools/utils_geo.py:88-117)."""
    R = np.asarray(R, dtype=np.float64)
    m = np.swapaxes(R, -1, -2)
    out = np.zeros(R.shape[:-2] + (4,))
    flat_m = m.reshape(-1, 3, 3)
    flat_o = out.reshape(-1, 4)
    for i, a in enumerate(flat_m):
        if a[2, 2] < 0:
            if a[0, 0] > a[1, 1]:
                t = 1 + a[0, 0] - a[1, 1] - a[2, 2]
                q = [a[1, 2] - a[2, 1], t, a[0, 1] + a[1, 0], a[2, 0] + a[0, 2]]
            else:
                t = 1 - a[0, 0] + a[1, 1] - a[2, 2]
                q = [a[2, 0] - a[0, 2], a[0, 1] + a[1, 0], t, a[1, 2] + a[2, 1]]
        else:
            if a[0, 0] < -a[1, 1]:
                t = 1 - a[0, 0] - a[1, 1] + a[2, 2]
                q = [a[0, 1] - a[1, 0], a[2, 0] + a[0, 2], a[1, 2] + a[2, 1], t]
            else:
                t = 1 + a[0, 0] + a[1, 1] + a[2, 2]
                q = [t, a[1, 2] - a[2, 1], a[2, 0] - a[0, 2], a[0, 1] - a[1, 0]]
        q = np.asarray(q) * (0.5 / np.sqrt(t))
        flat_o[i] = -q if q[0] < 0 else q
    return out


def virtual_grid(image_size=KITTI_IMAGE_SIZE, step: float = 0.1) -> np.ndarray:
    """10x10 grid of pixel positions, [100,2] (x,y)."""
    xx, yy = np.meshgrid(np.arange(0, 1, step), np.arange(0, 1, step))
    return np.stack([image_size[1] * xx.ravel(), image_size[0] * yy.ravel()], axis=1)


def make_batch(B: int, N: int, seed: int = 0, *, outlier_frac: float = 0.3,
               noise_px: float = 0.5, planar: bool = False,
               weight_mode: str = "softmax", K: np.ndarray = KITTI_K,
               image_size=KITTI_IMAGE_SIZE) -> dict:
    """One synthetic batch.  Keys mirror the reference's sample / data_batch dicts
    (deepFEPE/Train_model_pipeline.py:433-446):

    matches_xy_ori [B,N,4] f32 pixels (x1,y1,x2,y2); weights [B,1,N] f32 (sum 1 per pair);
    Ks, K_invs [B,3,3]; F_gt, E_gt [B,3,3]; delta_Rtijs_4_4 [B,4,4] scene motion;
    q_cam [B,4,1], t_cam [B,3,1] (inverse motion); pts1_virt, pts2_virt [B,100,3] homogeneous
    pixels on the GT epipolar geometry; inlier_mask [B,N] bool.
    """
    rng = np.random.default_rng(seed)
    H, W = image_size[0], image_size[1]

    rvec = rng.normal(0.0, 0.02, size=(B, 3))
    R = rodrigues(rvec)
    t = rng.normal(size=(B, 3)) * np.array([0.1, 0.05, 1.0])
    t /= np.linalg.norm(t, axis=1, keepdims=True)

    X = np.empty((B, N, 3))
    X[..., 0] = rng.uniform(-20, 20, size=(B, N))
    X[..., 1] = rng.uniform(-4, 4, size=(B, N))
    X[..., 2] = rng.uniform(4, 54, size=(B, N))
    if planar:  # all points on one plane: F is not unique
        nrm = np.array([0.05, -0.1, 1.0])
        X[..., 2] = (30.0 - X[..., 0] * nrm[0] - X[..., 1] * nrm[1]) / nrm[2]

    X2 = X @ np.swapaxes(R, 1, 2) + t[:, None, :]
    p1 = X @ K.T
    p2 = X2 @ K.T
    x1 = p1[..., :2] / p1[..., 2:3]
    x2 = p2[..., :2] / p2[..., 2:3]
    x1 = x1 + rng.normal(0.0, noise_px, size=x1.shape)
    x2 = x2 + rng.normal(0.0, noise_px, size=x2.shape)

    n_out = int(round(outlier_frac * N))
    inlier = np.ones((B, N), dtype=bool)
    if n_out > 0:
        x2[:, :n_out, 0] = rng.uniform(0, W, size=(B, n_out))
        x2[:, :n_out, 1] = rng.uniform(0, H, size=(B, n_out))
        inlier[:, :n_out] = False

    if weight_mode == "uniform":
        logits = np.zeros((B, N))
    elif weight_mode == "softmax":
        logits = rng.normal(size=(B, N))
    elif weight_mode == "peaked":          # softmax(3 * randn): few points dominate
        logits = 3.0 * rng.normal(size=(B, N))
    elif weight_mode == "inlier":          # inlier favouring: -4 on the outliers
        logits = rng.normal(size=(B, N)) * 0.5
        logits[~inlier] -= 4.0
    else:
        raise ValueError(f"unknown weight_mode {weight_mode!r}")
    logits = logits - logits.max(axis=1, keepdims=True)
    w = np.exp(logits)
    w /= w.sum(axis=1, keepdims=True)

    Kinv = np.linalg.inv(K)
    E = skew(t) @ R
    F = Kinv.T @ E @ Kinv

    # virtual correspondences on the GT geometry
    g = virtual_grid(image_size)
    g1 = np.concatenate([g, np.ones((g.shape[0], 1))], axis=1)          # [100,3]
    lines = g1 @ np.swapaxes(F, 1, 2)                                     # l2 = F x1, [B,100,3]
    a, b, c = lines[..., 0], lines[..., 1], lines[..., 2]
    d = (a * g[None, :, 0] + b * g[None, :, 1] + c) / (a * a + b * b + 1e-300)
    v2 = np.stack([g[None, :, 0] - a * d, g[None, :, 1] - b * d, np.ones_like(d)], axis=-1)
    v1 = np.broadcast_to(g1, v2.shape).copy()

    Rt = np.zeros((B, 4, 4))
    Rt[:, :3, :3] = R
    Rt[:, :3, 3] = t
    Rt[:, 3, 3] = 1.0
    Rt_cam = np.linalg.inv(Rt)
    q_cam = rot_to_quat(Rt_cam[:, :3, :3])[..., None]
    t_cam = Rt_cam[:, :3, 3:4]

    f32 = np.float32
    Kb = np.broadcast_to(K, (B, 3, 3)).astype(f32)
    return {
        "matches_xy_ori": np.concatenate([x1, x2], axis=2).astype(f32),
        "weights": w[:, None, :].astype(f32),
        "Ks": Kb.copy(),
        "K_invs": np.broadcast_to(Kinv, (B, 3, 3)).astype(f32).copy(),
        "F_gt": F.astype(f32),
        "E_gt": E.astype(f32),
        "delta_Rtijs_4_4": Rt.astype(f32),
        "q_cam": q_cam.astype(f32),
        "t_cam": t_cam.astype(f32),
        "pts1_virt": v1.astype(f32),
        "pts2_virt": v2.astype(f32),
        "inlier_mask": inlier,
        "matches_good_unique_nums": np.full((B,), N, dtype=np.int64),
        "image_size": tuple(image_size),
    }


def norm_hw_transform(image_size=KITTI_IMAGE_SIZE) -> np.ndarray:
    """The reference's NormalizeAndExpand_HW matrix (deepFEPE/models/DeepFNet.py:111)."""
    H, W = image_size[0], image_size[1]
    return np.array([[2.0 / W, 0, -1.0], [0, 2.0 / H, -1.0], [0, 0, 1.0]], dtype=np.float64)

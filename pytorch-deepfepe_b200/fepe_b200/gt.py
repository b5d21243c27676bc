"""Batched, on-device mirror of the ground-truth keys the reference's dataset builds per sample on the host.

Reference (numpy + OpenCV, one sample at a time inside the DataLoader workers):
    deepFEPE/datasets/kitti_odo_corr.py:290-302 get_E_F  -> dsac_tools/utils_F.py:835-846 E_F_from_Rt_np
    deepFEPE/datasets/kitti_odo_corr.py:526-541          -> dsac_tools/utils_misc.py:173-199 get_virt_x1x2_np
                                                            (cv2.correctMatches of the 10x10 grid of :163-171)
    deepFEPE/datasets/kitti_odo_corr.py:547-566          -> dsac_tools/utils_geo.py:88-117 R_to_q_np
Here: one launch of fepe_gt_virt for the whole batch; the results are born on the device, where the loss
(train_good_utils.py:325-326, :340-342) and the pose head read them.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops


def get_virt_x1x2_grid(im_shape, step: float = 0.1, device=None):
    """utils_misc.py:163-171: (pts1_virt_b, pts2_virt_b), two equal [100,2] float32 pixel grids."""
    xx, yy = np.meshgrid(np.arange(0, 1, step), np.arange(0, 1, step))
    g = np.float32(np.vstack((im_shape[1] * xx.flatten(), im_shape[0] * yy.flatten())).T)
    t = torch.from_numpy(g)
    if device is not None:
        t = t.to(device)
    return t, t.clone()


def get_virt_x1x2_batch(F_gt: torch.Tensor, K: torch.Tensor, pts1_virt_b: torch.Tensor, pts2_virt_b: torch.Tensor):
    """utils_misc.py:173-199 get_virt_x1x2_np for a batch: F_gt, K [B,3,3]; grids [P,2].
    Returns (pts1_virt_normalized, pts2_virt_normalized, pts1_virt, pts2_virt), each [B,P,3]; like the reference,
    both normalised tensors are K^-1 pts1_virt (:197-198)."""
    _, p1, p2, pn = ops.gt_virt(K, None, pts1_virt_b, pts2_virt_b, F_in=F_gt)
    return pn, pn.clone(), p1, p2


def gt_sample_batch(delta_Rtijs_4_4: torch.Tensor, Ks: torch.Tensor, image_size,
                    grids: Optional[tuple] = None) -> dict:
    """Every ground-truth key of the reference's sample dict that derives from (scene motion, K), for a batch:
    E, F [B,3,3]; pts1_virt, pts2_virt, pts1_virt_normalized, pts2_virt_normalized [B,100,3]; q_cam, q_scene [B,4,1];
    t_cam, t_scene [B,3,1] -- the shapes the collated reference batch has."""
    g1, g2 = grids if grids is not None else get_virt_x1x2_grid(image_size, device=Ks.device)
    gt, p1, p2, pn = ops.gt_virt(Ks, delta_Rtijs_4_4, g1, g2)
    B = gt.shape[0]
    return {"E": gt[:, 0:9].reshape(B, 3, 3), "F": gt[:, 9:18].reshape(B, 3, 3),
            "q_cam": gt[:, 18:22].reshape(B, 4, 1), "t_cam": gt[:, 22:25].reshape(B, 3, 1),
            "q_scene": gt[:, 25:29].reshape(B, 4, 1), "t_scene": gt[:, 29:32].reshape(B, 3, 1),
            "pts1_virt": p1, "pts2_virt": p2, "pts1_virt_normalized": pn, "pts2_virt_normalized": pn.clone()}

"""Batched, on-device mirror of the reference's per-sample validation pose recovery.

Reference (host, one sample at a time, fanned out over a pebble process pool):
    deepFEPE/Train_model_pipeline.py:954-964,1048-1061 -> train_good_utils.py:553-646 val_rt
    -> dsac_tools/utils_F.py:909-954 goodCorr_eval_nondecompose(p1s, p2s, E_hat, delta_Rtij_inv, K, scores)
       = cv2.recoverPose(E_hat, p1s, p2s, focal=K[0,0], pp=(K[0,2],K[1,2])) + utils_geo angular errors.
Here: one launch of fepe_recover_pose for the whole batch (and all layers); nothing leaves the device except the
[B,24] result rows the caller asks for.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def goodCorr_eval_nondecompose_batch(matches_xy: torch.Tensor, E_hat: torch.Tensor, delta_Rtijs_4_4: torch.Tensor,
                                     Ks: torch.Tensor, n_valid: Optional[torch.Tensor] = None, want_mask: bool = False):
    """matches_xy [B,N,4] pixels (x1,y1,x2,y2); E_hat [B,3,3] or [L,B,3,3]; delta_Rtijs_4_4 [B,4,4] (the scene motion
    the reference inverts at train_good_utils.py:581); Ks [B,3,3].

    Returns a dict: 'M' [L,B,3,4] = [R|t] (the reference's np.hstack((R, t))), 'err_q', 'err_t' [L,B] degrees,
    'num_inlier' [L,B] (cv2.recoverPose's return value), 'mask' [L,B,N] uint8 or None.  The reference's `scores`
    (top-10 % filter) is None at every call site (val_rt passes None) and is not provided."""
    out, mask = ops.recover_pose(E_hat, Ks, matches_xy, delta_Rtijs_4_4, n_valid=n_valid, want_mask=want_mask)
    L, B = out.shape[0], out.shape[1]
    M = torch.cat((out[..., :9].reshape(L, B, 3, 3), out[..., 9:12].reshape(L, B, 3, 1)), 3)
    return {"M": M, "err_q": out[..., 18], "err_t": out[..., 19], "num_inlier": out[..., 12].to(torch.int32),
            "counts": out[..., 14:18].to(torch.int32), "mask": mask}

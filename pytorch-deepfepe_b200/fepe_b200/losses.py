"""Loss-side mirrors of the reference's train_good_utils for the pose path, on the device.

`get_Rt_loss` keeps the reference's name, argument order and returned dict
(deepFEPE/train_good_utils.py:64-295) so that Train_model_pipeline.py:560-586 can call it unchanged:

    Rt_loss = get_Rt_loss(E_ests_layers, Ks_cpu, x1_cpu, x2_cpu, delta_Rtijs_4_4_cpu, qs_cam, ts_cam, device=...)
    loss = clamp(stack(Rt_loss['q_l2_error_layers_list']), 0, cq).mean() * bq + clamp(stack(.. 't_l2 ..'), 0, ct).mean() * bt

The reference walks layers x samples in Python, moves every E to the CPU (`E_ests.cpu()`, :106), decomposes it with a
LAPACK 3x3 SVD per sample (utils_F._get_M2s, utils_F.py:478-498), converts rotations to quaternions
(utils_geo._R_to_q), picks the nearer candidate (:160-188) and evaluates the angular metrics in numpy.  Here all of it
is ONE launch of fepe_pose_fwd for every (layer, sample) and ONE of fepe_pose_bwd in the backward pass
(ops.PoseLossFunction); the numpy metric arrays the caller logs cost a single device-to-host copy.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import ops


def get_Rt_loss(E_ests_layers: Sequence[torch.Tensor], Ks_cpu, x1_cpu, x2_cpu, delta_Rtijs_4_4_cpu: torch.Tensor,
                qs_cam: torch.Tensor, ts_cam: torch.Tensor, device="cuda", metrics_on_host: bool = True):
    """E_ests_layers: list (depth) of [B,3,3] CUDA tensors WITH autograd history; delta_Rtijs_4_4_cpu [B,4,4];
    qs_cam [B,4(,1)], ts_cam [B,3(,1)].  Ks_cpu, x1_cpu, x2_cpu are accepted and unused, exactly as in the reference
    (its docstring says "no use"; they only fed commented-out code).  Returns the reference's dict:
      t_l2_error_mean, q_l2_error_mean           scalars (differentiable)
      t_l2_error_list, q_l2_error_list           [depth] per-layer means -- the reference fills BOTH from the
                                                 translation list (train_good_utils.py:269-270); kept
      R_angle_error_mean / _list, t_angle_error_mean / _list      floats / numpy [depth]
      R_angle_error_layers_list, t_angle_error_layers_list        list of numpy [B]
      t_l2_error_layers_list, q_l2_error_layers_list              list of [B] tensors (differentiable)
    metrics_on_host=False keeps the angular metrics on the device (CUDA tensors under the same keys, no device-to-host
    copy inside the call) -- for callers that capture the step in a CUDA graph and fetch the metrics afterwards.
    """
    E = torch.stack(list(E_ests_layers))                          # [L,B,3,3]
    if not E.is_cuda:
        raise RuntimeError("fepe_b200.get_Rt_loss needs CUDA tensors: there is no CPU path")
    L, B = E.shape[0], E.shape[1]
    dev = E.device
    eye = torch.eye(3, device=dev, dtype=torch.float32).expand(B, 3, 3).contiguous()
    Rt = delta_Rtijs_4_4_cpu.to(dev, torch.float32)
    q = qs_cam.to(dev, torch.float32).reshape(B, 4)
    t = ts_cam.to(dev, torch.float32).reshape(B, 3)
    # K = I and the identity affine make the head's E = (TK)^T F (TK) equal to its input
    q_l2, t_l2, _, out = ops.PoseLossFunction.apply(E.float(), eye, q, t, Rt, None, None, *ops.IDENTITY_AFFINE, 0.02)
    if metrics_on_host:
        ang = out[..., 23:25].detach().cpu().numpy().astype(np.float64)     # the one D2H copy: [L,B,2]
    else:
        ang = out[..., 23:25].detach()
    R_ang, t_ang = ang[..., 0], ang[..., 1]
    t_l2_mean_layers = t_l2.mean(1)
    q_l2_mean_layers = q_l2.mean(1)
    return {
        "t_l2_error_mean": t_l2_mean_layers.mean(),
        "q_l2_error_mean": q_l2_mean_layers.mean(),
        "t_l2_error_list": t_l2_mean_layers,
        "q_l2_error_list": t_l2_mean_layers,                      # sic (train_good_utils.py:270)
        "R_angle_error_mean": float(R_ang.mean(1).mean()) if metrics_on_host else R_ang.mean(),
        "R_angle_error_list": R_ang.mean(1),
        "t_angle_error_mean": float(t_ang.mean(1).mean()) if metrics_on_host else t_ang.mean(),
        "t_angle_error_list": t_ang.mean(1),
        "R_angle_error_layers_list": [R_ang[l] for l in range(L)],
        "t_angle_error_layers_list": [t_ang[l] for l in range(L)],
        "t_l2_error_layers_list": [t_l2[l] for l in range(L)],
        "q_l2_error_layers_list": [q_l2[l] for l in range(L)],
    }


def pose_loss_from_Rt_loss(Rt_loss: dict, clamp_q: float = 0.1, clamp_t: float = 0.5, balance_q: float = 1.0,
                           balance_t: float = 0.1) -> torch.Tensor:
    """Train_model_pipeline.py:580-586 (first stage of the clamp schedule :474-489, balance of
    configs/kitti_corr_baseline.yaml:50-51 by default)."""
    q = torch.stack(Rt_loss["q_l2_error_layers_list"])
    t = torch.stack(Rt_loss["t_l2_error_layers_list"])
    return torch.clamp(q, 0.0, clamp_q).mean() * balance_q + torch.clamp(t, 0.0, clamp_t).mean() * balance_t


def deepf_training_loss(out_layers: Sequence[torch.Tensor], Ks: torch.Tensor, pts1_virt: torch.Tensor,
                        pts2_virt: torch.Tensor, delta_Rtijs_4_4: torch.Tensor, qs_cam: torch.Tensor,
                        ts_cam: torch.Tensor, affine, clamp_at: float = 0.02, clamp_q: float = 0.1,
                        clamp_t: float = 0.5, balance_q: float = 1.0, balance_t: float = 0.1):
    """The whole loss of a DeepF training step in ONE launch (and one in the backward pass):

        loss_F   = mean over layers of mean(compute_epi_residual(T1 virt1, T1 virt2, F_i, clamp_at))
                                                        -- get_all_loss_DeepF, train_good_utils.py:325-364
        E_i      = K^T T2^T F_i T1 K                    -- :356-358
        q/t loss = get_Rt_loss(E_layers, ...)           -- :64-295, combined as Train_model_pipeline.py:580-592

    instead of the reference's (and `get_Rt_loss`'s + torch glue's) ~25 small kernels per layer: the head computes E
    from F, K and the image affine itself (ops.PoseLossFunction / fepe_pose_fwd over every (layer, pair)).
    out_layers: list (depth) of [B,3,3] (DeepFNet's `out_layers`); pts*_virt [B,V,3] homogeneous PIXEL coordinates;
    affine = ops.hw_affine(image_size).  Returns (loss, dict) -- the dict holds per-(layer, pair) tensors on the device:
    q_l2, t_l2, loss_F [L,B] (differentiable) and R_angle, t_angle [L,B] (metrics)."""
    F = torch.stack(list(out_layers))
    if not F.is_cuda:
        raise RuntimeError("fepe_b200.deepf_training_loss needs CUDA tensors: there is no CPU path")
    B = F.shape[1]
    dev = F.device
    q = qs_cam.to(dev, torch.float32).reshape(B, 4)
    t = ts_cam.to(dev, torch.float32).reshape(B, 3)
    Rt = delta_Rtijs_4_4.to(dev, torch.float32)
    q_l2, t_l2, loss_F, out = ops.PoseLossFunction.apply(F.float(), Ks.to(dev, torch.float32), q, t, Rt,
                                                         pts1_virt.contiguous(), pts2_virt.contiguous(), *affine, clamp_at)
    loss = (loss_F.mean() + torch.clamp(q_l2, 0.0, clamp_q).mean() * balance_q
            + torch.clamp(t_l2, 0.0, clamp_t).mean() * balance_t)
    return loss, {"q_l2": q_l2, "t_l2": t_l2, "loss_F": loss_F, "R_angle": out[..., 23], "t_angle": out[..., 24]}

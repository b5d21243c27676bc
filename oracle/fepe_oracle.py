"""CPU ORACLE for the deepFEPE hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, op for op, the reference's differentiable weighted 8-point /
relative-pose path on CPU torch tensors.  It is the checker for the CUDA kernels:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it.  The product package
(``pytorch-deepfepe_b200/``) never does and has no CPU fallback.

PARITY PIN: the reference (eric-yyjau/pytorch-deepFEPE @ 7f3e775) holds no golden vectors
or numerical tests for this path, so this oracle is pinned against
outputs of the reference's own modules executed in the build container:
``tests/golden/make_golden.py`` imports the unmodified reference from /root/reference
(with import stubs for matplotlib/pebble/superpoint), runs ``Fit``, ``NormalizeAndExpand_HW``,
``compute_epi_residual``, ``get_all_loss_DeepF``, ``_get_M2s``, ``_R_to_q``, ``get_Rt_loss``
and ``ErrorEstimator`` on seeded inputs and commits the results as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them.

Everything is dtype generic: fp32 reproduces the reference, fp64 serves as "truth" when
the CUDA result and the fp32 reference disagree at the 1e-6 level.

Reference map (paths under /root/reference/deepFEPE/):
  norm_hw                 models/DeepFNet.py:93-120    NormalizeAndExpand_HW
  hartley                 models/DeepFNet.py:148-179   Fit.normalize (called with ones, :194-199)
  fit_weighted_svd        models/DeepFNet.py:181-257   Fit.weighted_svd
  epi_residual            dsac_tools/utils_F.py:400-413 compute_epi_residual
  f_loss_layers           train_good_utils.py:325-369  get_all_loss_DeepF (F-loss + E per layer)
  essential_decompose     dsac_tools/utils_F.py:478-498 _get_M2s
  rot_to_quat             dsac_tools/utils_geo.py:58-86 _R_to_q
  pose_errors             train_good_utils.py:96-188   get_Rt_loss inner loop
  pose_loss               Train_model_pipeline.py:580-586 clamp / mean / balance
  build_error_estimator   models/ErrorEstimators.py:46-64
  deepf_forward           models/DeepFNet.py:429-554   DeepFNet.forward
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F_


# ----------------------------------------------------------------------------- a1
def norm_hw(matches_xy: torch.Tensor, image_size: Sequence[int]):
    """pixels [B,N,4] -> pts1, pts2 [B,N,3] in [-1,1]^2 (homogeneous), T [B,3,3].

    DeepFNet.py:108-120 builds T = [[2/W,0,-1],[0,2/H,-1],[0,0,1]] and applies it to both
    images; DeepFNet.get_input (:377-383) permutes the result back to [B,N,3]."""
    H, W = image_size[0], image_size[1]
    B, N, _ = matches_xy.shape
    T = torch.tensor([[2.0 / W, 0.0, -1.0], [0.0, 2.0 / H, -1.0], [0.0, 0.0, 1.0]],
                     dtype=matches_xy.dtype, device=matches_xy.device)
    ones = torch.ones(B, N, 1, dtype=matches_xy.dtype, device=matches_xy.device)
    h1 = torch.cat((matches_xy[:, :, :2], ones), 2)
    h2 = torch.cat((matches_xy[:, :, 2:], ones), 2)
    Tb = T.unsqueeze(0).expand(B, -1, -1)
    pts1 = (Tb @ h1.transpose(1, 2)).transpose(1, 2)
    pts2 = (Tb @ h2.transpose(1, 2)).transpose(1, 2)
    return pts1, pts2, Tb


# ----------------------------------------------------------------------------- a5
def hartley(pts: torch.Tensor):
    """Fit.normalize with unit weights (DeepFNet.py:148-179, weights=ones :194-199).

    pts [B,N,3] -> (ptsn [B,3,N], T [B,3,3]); scale uses the literal 1.4142."""
    B, N, _ = pts.shape
    c = pts.sum(1) / N                                           # [B,3]
    centred = pts - c.unsqueeze(1)
    meandist = centred[:, :, :2].pow(2).sum(2).sqrt().sum(1) / N  # [B]
    scale = 1.4142 / meandist
    T = torch.zeros(B, 3, 3, dtype=pts.dtype, device=pts.device)
    T[:, 0, 0] = scale
    T[:, 1, 1] = scale
    T[:, 2, 2] = 1
    T[:, 0, 2] = -c[:, 0] * scale
    T[:, 1, 2] = -c[:, 1] * scale
    return torch.bmm(T, pts.transpose(1, 2)), T


# ----------------------------------------------------------------------------- a6
def constraint_rows(pts1n: torch.Tensor, pts2n: torch.Tensor) -> torch.Tensor:
    """Row i = [x2*x1, x2*y1, x2, y2*x1, y2*y1, y2, x1, y1, 1] (DeepFNet.py:203-205), L2
    normalised per row (F.normalize, :211-212).  Inputs are [B,3,N]; output [B,N,9]."""
    p = torch.cat((pts2n[:, 0:1] * pts1n, pts2n[:, 1:2] * pts1n, pts1n), 1).transpose(1, 2)
    return F_.normalize(p, dim=2)


def fit_weighted_svd(pts1: torch.Tensor, pts2: torch.Tensor, weights: torch.Tensor,
                     canonical_sign: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fit.forward / weighted_svd (DeepFNet.py:181-257).

    pts1, pts2 [B,N,3] homogeneous, weights [B,1,N].  Returns out [B,3,3] (de-normalised,
    rank-2) and the signed residual X @ f [B,N].  The per-pair torch.svd loop of the
    reference (:232-240) is kept: the smallest right singular vector of X = w * p_hat gives
    f; the 3x3 SVD zeroes the last singular value; out = T2^T F_ T1 (:256).

    canonical_sign: LAPACK's sign of V[:,-1] is arbitrary; with True the vector is flipped so that its
    largest-magnitude entry is positive -- the convention of the CUDA kernel (include/fepe_b200.h: fepe_fit_fwd).
    Any sign is a valid SVD; F and the signed residual follow it."""
    w = weights.squeeze(1).unsqueeze(2)                          # [B,N,1]
    pts1n, T1 = hartley(pts1)
    pts2n, T2 = hartley(pts2)
    X = constraint_rows(pts1n, pts2n) * w                        # weights enter un-squared (:214)
    mask = torch.tensor([1.0, 1.0, 0.0], dtype=X.dtype, device=X.device)
    Fs, fvecs = [], []
    for b in range(X.shape[0]):
        _, _, V = torch.svd(X[b])
        v = V[:, -1]
        if canonical_sign:
            v = v * torch.sign(v.detach()[v.detach().abs().argmax()])
        fvecs.append(v / v.norm())
        U3, S3, V3 = torch.svd(v.view(3, 3))
        Fs.append(U3 @ torch.diag(S3 * mask) @ V3.t())
    Fn = torch.stack(Fs)
    fv = torch.stack(fvecs)
    residual = (X @ fv.unsqueeze(-1)).squeeze(-1)
    out = T2.transpose(1, 2) @ Fn @ T1
    return out, residual


# ----------------------------------------------------------------------------- a7
def epi_residual(pts1: torch.Tensor, pts2: torch.Tensor, Fm: torch.Tensor,
                 clamp_at: Optional[float] = 0.5) -> torch.Tensor:
    """compute_epi_residual (utils_F.py:400-413): |x2^T F x1| (1/(|l1_xy|+1e-6)+1/(|l2_xy|+1e-6)),
    clamped from above.  pts [B,N,3], F [B,3,3] -> [B,N]."""
    l1 = pts2 @ Fm                       # rows x2^T F
    l2 = pts1 @ Fm.transpose(1, 2)       # rows (F x1)^T
    dd = (pts1 * l1).sum(2)
    d = dd.abs() * (1.0 / (l1[:, :, :2].norm(2, 2) + 1e-6) + 1.0 / (l2[:, :, :2].norm(2, 2) + 1e-6))
    return d if clamp_at is None else torch.clamp(d, max=clamp_at)


# ----------------------------------------------------------------------------- a9
def f_loss_layers(out_layers: Sequence[torch.Tensor], T1: torch.Tensor, T2: torch.Tensor,
                  pts1_virt: torch.Tensor, pts2_virt: torch.Tensor, Ks: torch.Tensor,
                  clamp_at: float = 0.02):
    """The F-loss part of get_all_loss_DeepF (train_good_utils.py:325-369).

    Returns (loss_F, per-layer [B,100] losses, E_ests_layers)."""
    p1 = (T1 @ pts1_virt.transpose(1, 2)).transpose(1, 2)
    p2 = (T2 @ pts2_virt.transpose(1, 2)).transpose(1, 2)
    losses, E_layers, total = [], [], 0.0
    for Fo in out_layers:
        l = epi_residual(p1, p2, Fo, clamp_at)
        losses.append(l)
        total = total + l.mean()
        E_layers.append(Ks.transpose(1, 2) @ T2.transpose(1, 2) @ Fo @ T1 @ Ks)
    return total / len(out_layers), losses, E_layers


# ----------------------------------------------------------------------------- a11
def essential_decompose(E: torch.Tensor):
    """_get_M2s (utils_F.py:478-498) for one 3x3 E: returns ([R1,R2],[t,-t])."""
    U, S, V = torch.svd(E)
    W = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], dtype=E.dtype)
    if torch.det(U @ W @ V.t()) < 0:
        W = -W
    t = U[:, 2:3] / torch.norm(U[:, 2:3])
    return [U @ W @ V.t(), U @ W.t() @ V.t()], [t, -t]


# ----------------------------------------------------------------------------- a12
def rot_to_quat(R: torch.Tensor) -> torch.Tensor:
    """_R_to_q (utils_geo.py:58-86): 4-branch trace method on m = R^T, w>=0, returns [4,1]."""
    m = R.t()
    if m[2, 2] < 0:
        if m[0, 0] > m[1, 1]:
            t = 1 + m[0, 0] - m[1, 1] - m[2, 2]
            q = torch.stack((m[1, 2] - m[2, 1], t, m[0, 1] + m[1, 0], m[2, 0] + m[0, 2]))
        else:
            t = 1 - m[0, 0] + m[1, 1] - m[2, 2]
            q = torch.stack((m[2, 0] - m[0, 2], m[0, 1] + m[1, 0], t, m[1, 2] + m[2, 1]))
    else:
        if m[0, 0] < -m[1, 1]:
            t = 1 - m[0, 0] - m[1, 1] + m[2, 2]
            q = torch.stack((m[0, 1] - m[1, 0], m[2, 0] + m[0, 2], m[1, 2] + m[2, 1], t))
        else:
            t = 1 + m[0, 0] + m[1, 1] + m[2, 2]
            q = torch.stack((t, m[1, 2] - m[2, 1], m[2, 0] - m[0, 2], m[0, 1] - m[1, 0]))
    q = q * (0.5 / torch.sqrt(t))
    if q[0] < 0:
        q = -q
    return q.unsqueeze(-1)


# ----------------------------------------------------------------------------- a10
def pose_errors(E_ests: torch.Tensor, q_cam: torch.Tensor, t_cam: torch.Tensor,
                Rt_scene: torch.Tensor):
    """Inner loop of get_Rt_loss for ONE layer (train_good_utils.py:96-188).

    E_ests [B,3,3]; q_cam [B,4,1]; t_cam [B,3,1]; Rt_scene [B,4,4] (delta_Rtijs_4_4).
    Per pair: decompose E^T (:106), quaternions of both rotations, L2 to the GT quaternion and
    to the normalised GT translation, independent min-select for q and t (:160-188).
    Returns q_l2 [B], t_l2 [B] (differentiable) and R_angle_deg [B], t_angle_deg [B]."""
    B = E_ests.shape[0]
    Rt_inv = torch.inverse(Rt_scene)
    q_l2, t_l2, r_ang, t_ang = [], [], [], []
    for b in range(B):
        Rs, ts = essential_decompose(E_ests[b].t())
        qa, qb = rot_to_quat(Rs[0]), rot_to_quat(Rs[1])
        tg = F_.normalize(t_cam[b], p=2, dim=0)
        eq = [torch.norm(qa - q_cam[b]), torch.norm(qb - q_cam[b])]
        et = [torch.norm(ts[0] - tg), torch.norm(ts[1] - tg)]
        iq = 0 if bool(eq[0] < eq[1]) else 1
        it = 0 if bool(et[0] < et[1]) else 1
        q_l2.append(eq[iq])
        t_l2.append(et[it])
        # metrics: rotation angle of R_est R_gt^T (cv2.Rodrigues norm, utils_geo.py:150-152,
        # equals acos((tr-1)/2)) and the angle between translations (:175-179)
        # cv2.Rodrigues first projects onto SO(3); atan2(|antisym part|, (tr-1)/2) gives the
        # same angle without needing that projection and stays accurate for tiny rotations.
        Rd = Rs[iq].detach().double() @ Rt_inv[b, :3, :3].double().t()
        cosr = float((torch.trace(Rd) - 1.0) / 2.0)
        sinr = 0.5 * math.sqrt(float((Rd[2, 1] - Rd[1, 2]) ** 2 + (Rd[0, 2] - Rd[2, 0]) ** 2
                                     + (Rd[1, 0] - Rd[0, 1]) ** 2))
        r_ang.append(math.degrees(math.atan2(sinr, cosr)))
        a, c = ts[it].detach().double().flatten(), tg.detach().double().flatten()
        cost = float((a @ c) / ((a.norm() + 1e-10) * (c.norm() + 1e-10) + 1e-10))
        t_ang.append(math.degrees(math.acos(max(-1.0, min(1.0, cost)))))
    return (torch.stack(q_l2), torch.stack(t_l2),
            torch.tensor(r_ang, dtype=torch.float64), torch.tensor(t_ang, dtype=torch.float64))


def pose_loss(E_layers: Sequence[torch.Tensor], q_cam, t_cam, Rt_scene,
              clamp_q: float = 0.1, clamp_t: float = 0.5,
              balance_q: float = 1.0, balance_t: float = 0.1):
    """loss = clamp(q_l2,0,cq).mean()*bq + clamp(t_l2,0,ct).mean()*bt over layers x batch
    (Train_model_pipeline.py:580-586; defaults = first stage of the clamp schedule :474-489 and
    balance_q/t of configs/kitti_corr_baseline.yaml:50-51)."""
    qs, ts, ra, ta = [], [], [], []
    for E in E_layers:
        q, t, r, a = pose_errors(E, q_cam, t_cam, Rt_scene)
        qs.append(q), ts.append(t), ra.append(r), ta.append(a)
    q_all, t_all = torch.stack(qs), torch.stack(ts)
    loss = torch.clamp(q_all, 0.0, clamp_q).mean() * balance_q + \
        torch.clamp(t_all, 0.0, clamp_t).mean() * balance_t
    return loss, q_all, t_all, torch.stack(ra), torch.stack(ta)


# ----------------------------------------------------------------------------- a3
def build_error_estimator(input_size: int, output_size: int = 1) -> nn.Sequential:
    """ErrorEstimator live branch (ErrorEstimators.py:46-64): five (Conv1d k=1 -> InstanceNorm1d
    affine -> LeakyReLU 0.01) blocks 64/128/1024/512/256 and a final Conv1d to ``output_size``.
    Indices in the Sequential match the reference so state_dicts are interchangeable."""
    chans = [input_size, 64, 128, 1024, 512, 256]
    layers: List[nn.Module] = []
    for cin, cout in zip(chans[:-1], chans[1:]):
        layers += [nn.Conv1d(cin, cout, kernel_size=1, bias=True),
                   nn.InstanceNorm1d(cout, affine=True),
                   nn.LeakyReLU(inplace=False)]
    layers.append(nn.Conv1d(chans[-1], output_size, kernel_size=1, bias=True))
    return nn.Sequential(*layers)


# ----------------------------------------------------------------------------- a8
def deepf_forward(matches_xy: torch.Tensor, image_size, net_init: nn.Module, net_update: nn.Module,
                  depth: int = 5, quality: Optional[torch.Tensor] = None,
                  weights_im: Optional[torch.Tensor] = None, canonical_sign: bool = False) -> dict:
    """DeepFNet.forward (DeepFNet.py:429-554) for the live option set (no offsets / des / tri).

    Returns the same dict keys the reference returns (:534-548).  canonical_sign: see fit_weighted_svd -- the signed
    residual is an input channel of the next layer's network (:487), so layers >= 1 depend on the convention."""
    pts1, pts2, T = norm_hw(matches_xy, image_size)
    feats = [(pts1[:, :, :2] + 1) / 2, (pts2[:, :, :2] + 1) / 2]
    if quality is not None:
        feats.append(quality)
    net_in0 = torch.cat(feats, 2).transpose(1, 2)                # [B,4(+Q),N]
    logits = net_init(net_in0)
    w = F_.softmax(logits, dim=2)
    if weights_im is not None:
        w = w * weights_im
    out_layers, epi_layers, res_layers, w_layers, logit_layers = [], [], [], [w], [logits]
    for _ in range(depth - 1):
        out, res = fit_weighted_svd(pts1, pts2, w, canonical_sign)
        out_layers.append(out)
        res_layers.append(res)
        epi = epi_residual(pts1, pts2, out).unsqueeze(1)
        epi_layers.append(epi)
        logits = net_update(torch.cat((net_in0, w, epi, res.unsqueeze(1)), 1))
        w = F_.softmax(logits, dim=2)
        if weights_im is not None:
            w = w * weights_im
        w_layers.append(w)
        logit_layers.append(logits)
    out, res = fit_weighted_svd(pts1, pts2, w, canonical_sign)
    out_layers.append(out)
    res_layers.append(res)
    return {"logits": logits.squeeze(1), "logits_layers": logit_layers, "F_est": out,
            "epi_res_layers": epi_layers, "T1": T, "T2": T, "out_layers": out_layers,
            "pts1": pts1, "pts2": pts2, "weights": w, "residual_layers": res_layers,
            "weights_layers": w_layers}


# ----------------------------------------------------------------------------- helpers
def sign_aligned_rel_err(Fa: torch.Tensor, Fb: torch.Tensor) -> torch.Tensor:
    """min(|Fa-Fb|, |Fa+Fb|)_F / |Fb|_F per pair -- V[:,-1]'s sign is arbitrary."""
    a = Fa.reshape(Fa.shape[0], -1).double()
    b = Fb.reshape(Fb.shape[0], -1).double()
    return torch.minimum((a - b).norm(dim=1), (a + b).norm(dim=1)) / b.norm(dim=1)

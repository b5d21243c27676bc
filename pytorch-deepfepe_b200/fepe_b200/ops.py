"""Tensor-level entry points over the C ABI (device pointers + current CUDA stream).

PyTorch is used for device memory, streams and autograd plumbing only; all arithmetic of the
path happens in libfepe_b200.so.  Inputs must be CUDA fp32 tensors; anything else raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

IDENTITY_AFFINE = (1.0, 0.0, 1.0, 0.0)


def hw_affine(image_size) -> Tuple[float, float, float, float]:
    """(ax,bx,ay,by) of the reference's NormalizeAndExpand_HW (deepFEPE/models/DeepFNet.py:111)."""
    H, W = float(image_size[0]), float(image_size[1])
    return (2.0 / W, -1.0, 2.0 / H, -1.0)


def _check_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"fepe_b200: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"fepe_b200: {name} must be float32, got {t.dtype}")
    return t.contiguous()


def _stream_ptr(device=None) -> int:
    """Raw handle of the current stream of `device` (default: the current device; every caller sits inside a
    torch.cuda.device(...) guard of the tensors' device)."""
    return torch.cuda.current_stream(device).cuda_stream


def fit_forward(matches: torch.Tensor, weights: torch.Tensor, affine=IDENTITY_AFFINE,
                clamp_at: float = 0.5, want_epi: bool = True, want_saved: bool = False, out=None):
    """matches [B,N,4], weights [B,N] (or [B,1,N]) -> F [B,3,3], residual [B,N], epi [B,N]|None,
    saved [B,64] float64|None.  One launch of the fused kernel (include/fepe_b200.h: fepe_fit_fwd)."""
    matches = _check_cuda_f32(matches, "matches")
    weights = _check_cuda_f32(weights, "weights")
    B, N, four = matches.shape
    if four != 4:
        raise RuntimeError("fepe_b200: matches must be [B,N,4] (x1,y1,x2,y2)")
    weights = weights.reshape(B, N)
    dev = matches.device
    with torch.cuda.device(dev):
        if out is not None:      # caller-owned output buffers (F, res, epi|None, saved|None), e.g. for graphs
            F, res, epi, saved = out
        else:
            F = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
            res = torch.empty(B, N, dtype=torch.float32, device=dev)
            epi = torch.empty(B, N, dtype=torch.float32, device=dev) if want_epi else None
            saved = torch.empty(B, _lib.SAVED_DOUBLES, dtype=torch.float64, device=dev) if want_saved else None
        st = _lib.lib().fepe_fit_fwd(matches.data_ptr(), weights.data_ptr(), B, N,
                                     affine[0], affine[1], affine[2], affine[3], float(clamp_at),
                                     F.data_ptr(), res.data_ptr(),
                                     epi.data_ptr() if epi is not None else None,
                                     saved.data_ptr() if saved is not None else None,
                                     _stream_ptr())
    _lib.check(st, "fepe_fit_fwd")
    return F, res, epi, saved


def pose_forward(F: torch.Tensor, K: torch.Tensor, affine, q_gt: torch.Tensor, t_gt: torch.Tensor,
                 Rt_scene: Optional[torch.Tensor] = None, virt1: Optional[torch.Tensor] = None,
                 virt2: Optional[torch.Tensor] = None, clamp_at: float = 0.02,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F [L,B,3,3] or [B,3,3]; K [B,3,3]; q_gt [B,4(,1)]; t_gt [B,3(,1)]; Rt_scene [B,4,4];
    virt1/2 [B,V,3].  Returns [L,B,32] (layout in include/fepe_b200.h: fepe_pose_fwd)."""
    F = _check_cuda_f32(F, "F")
    if F.dim() == 3:
        F = F.unsqueeze(0)
    L, B = F.shape[0], F.shape[1]
    K = _check_cuda_f32(K, "K").reshape(B, 9)
    q_gt = _check_cuda_f32(q_gt, "q_gt").reshape(B, 4)
    t_gt = _check_cuda_f32(t_gt, "t_gt").reshape(B, 3)
    rt = _check_cuda_f32(Rt_scene, "Rt_scene").reshape(B, 16) if Rt_scene is not None else None
    V = 0
    if virt1 is not None:
        virt1 = _check_cuda_f32(virt1, "virt1")
        virt2 = _check_cuda_f32(virt2, "virt2")
        V = virt1.shape[1]
    with torch.cuda.device(F.device):
        if out is None:
            out = torch.empty(L, B, _lib.POSE_OUT_FLOATS, dtype=torch.float32, device=F.device)
        st = _lib.lib().fepe_pose_fwd(F.data_ptr(), K.data_ptr(), L, B, affine[0], affine[1], affine[2], affine[3],
                                      q_gt.data_ptr(), t_gt.data_ptr(), rt.data_ptr() if rt is not None else None,
                                      virt1.data_ptr() if virt1 is not None else None,
                                      virt2.data_ptr() if virt2 is not None else None, V, float(clamp_at),
                                      out.data_ptr(), _stream_ptr())
    _lib.check(st, "fepe_pose_fwd")
    return out


def fit_pose_forward(matches: torch.Tensor, weights: torch.Tensor, affine, K: torch.Tensor, q_gt: torch.Tensor,
                     t_gt: torch.Tensor, Rt_scene: Optional[torch.Tensor] = None, virt1: Optional[torch.Tensor] = None,
                     virt2: Optional[torch.Tensor] = None, clamp_at: float = 0.5, virt_clamp_at: float = 0.02,
                     want_epi: bool = True, want_saved: bool = False, out=None):
    """fit_forward + pose_forward (one layer) as ONE call of the C ABI (include/fepe_b200.h: fepe_fit_pose_fwd):
    a single kernel for batches of up to two pairs per SM.  Returns (F, residual, epi|None, saved|None, pose [B,32]);
    `out` may carry caller-owned buffers in that order."""
    matches = _check_cuda_f32(matches, "matches")
    weights = _check_cuda_f32(weights, "weights")
    B, N, four = matches.shape
    if four != 4:
        raise RuntimeError("fepe_b200: matches must be [B,N,4] (x1,y1,x2,y2)")
    weights = weights.reshape(B, N)
    dev = matches.device
    K = _check_cuda_f32(K, "K").reshape(B, 9)
    q_gt = _check_cuda_f32(q_gt, "q_gt").reshape(B, 4)
    t_gt = _check_cuda_f32(t_gt, "t_gt").reshape(B, 3)
    rt = _check_cuda_f32(Rt_scene, "Rt_scene").reshape(B, 16) if Rt_scene is not None else None
    V = 0
    if virt1 is not None:
        virt1 = _check_cuda_f32(virt1, "virt1")
        virt2 = _check_cuda_f32(virt2, "virt2")
        V = virt1.shape[1]
    with torch.cuda.device(dev):
        if out is not None:
            F, res, epi, saved, pose = out
        else:
            F = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
            res = torch.empty(B, N, dtype=torch.float32, device=dev)
            epi = torch.empty(B, N, dtype=torch.float32, device=dev) if want_epi else None
            saved = torch.empty(B, _lib.SAVED_DOUBLES, dtype=torch.float64, device=dev) if want_saved else None
            pose = torch.empty(B, _lib.POSE_OUT_FLOATS, dtype=torch.float32, device=dev)
        st = _lib.lib().fepe_fit_pose_fwd(matches.data_ptr(), weights.data_ptr(), B, N,
                                          affine[0], affine[1], affine[2], affine[3], float(clamp_at),
                                          F.data_ptr(), res.data_ptr(),
                                          epi.data_ptr() if epi is not None else None,
                                          saved.data_ptr() if saved is not None else None,
                                          K.data_ptr(), q_gt.data_ptr(), t_gt.data_ptr(),
                                          rt.data_ptr() if rt is not None else None,
                                          virt1.data_ptr() if virt1 is not None else None,
                                          virt2.data_ptr() if virt2 is not None else None, V, float(virt_clamp_at),
                                          pose.data_ptr(), _stream_ptr())
    _lib.check(st, "fepe_fit_pose_fwd")
    return F, res, epi, saved, pose


def fit_backward(matches: torch.Tensor, weights: torch.Tensor, saved: torch.Tensor, gF: torch.Tensor,
                 gres: Optional[torch.Tensor], gepi: Optional[torch.Tensor], affine=IDENTITY_AFFINE,
                 clamp_at: float = 0.5, want_coords: bool = False, out=None):
    """d loss / d weights [B,N] from the upstream gradients of (F, residual, epi).  One launch of
    fepe_fit_bwd (include/fepe_b200.h).  `want_coords`: also d loss / d matches [B,N,4] (fepe_fit_bwd_coords);
    returns (gweights, gmatches) then."""
    matches = _check_cuda_f32(matches, "matches")
    weights = _check_cuda_f32(weights, "weights")
    B, N, _ = matches.shape
    weights = weights.reshape(B, N)
    if saved is None or saved.dtype != torch.float64 or tuple(saved.shape) != (B, _lib.SAVED_DOUBLES):
        raise RuntimeError("fepe_b200: `saved` must be the [B,64] float64 state returned by fit_forward(want_saved=True)")
    gF = _check_cuda_f32(gF, "gF").reshape(B, 9)
    gres = _check_cuda_f32(gres, "gres").reshape(B, N) if gres is not None else None
    gepi = _check_cuda_f32(gepi, "gepi").reshape(B, N) if gepi is not None else None
    with torch.cuda.device(matches.device):
        if out is not None:          # caller-owned (gweights, gmatches | None)
            gw, gm = out
        else:
            gw = torch.empty(B, N, dtype=torch.float32, device=matches.device)
            gm = torch.empty(B, N, 4, dtype=torch.float32, device=matches.device) if want_coords else None
        st = _lib.lib().fepe_fit_bwd_coords(matches.data_ptr(), weights.data_ptr(), B, N,
                                            affine[0], affine[1], affine[2], affine[3], float(clamp_at),
                                            saved.contiguous().data_ptr(), gF.data_ptr(),
                                            gres.data_ptr() if gres is not None else None,
                                            gepi.data_ptr() if gepi is not None else None,
                                            gw.data_ptr(), gm.data_ptr() if gm is not None else None, _stream_ptr())
    _lib.check(st, "fepe_fit_bwd_coords")
    return (gw, gm) if want_coords else gw


def recover_pose(E: torch.Tensor, K: torch.Tensor, matches: torch.Tensor, Rt_scene: Optional[torch.Tensor] = None,
                 n_valid: Optional[torch.Tensor] = None, distance_thresh: float = 50.0, want_mask: bool = False):
    """Validation pose recovery on the device (include/fepe_b200.h: fepe_recover_pose) -- what
    cv2.recoverPose(E, p1, p2, focal=K[0,0], pp=(K[0,2],K[1,2])) + the angular errors against the ground truth do per
    sample in deepFEPE/dsac_tools/utils_F.py:909-954.

    E [L,B,3,3] or [B,3,3]; K [B,3,3]; matches [B,N,4] pixels; Rt_scene [B,4,4] (delta_Rtijs_4_4); n_valid [B] int32.
    Returns (out [L,B,24], mask [L,B,N] uint8 | None); out[..., :9] = R, [9:12] = t, [12] = points in front,
    [18] / [19] = err_q / err_t in degrees."""
    E = _check_cuda_f32(E, "E")
    if E.dim() == 3:
        E = E.unsqueeze(0)
    L, B = E.shape[0], E.shape[1]
    K = _check_cuda_f32(K, "K").reshape(B, 9)
    matches = _check_cuda_f32(matches, "matches")
    if matches.dim() != 3 or matches.shape[0] != B or matches.shape[2] != 4:
        raise RuntimeError("fepe_b200: matches must be [B,N,4] (x1,y1,x2,y2)")
    N = matches.shape[1]
    rt = _check_cuda_f32(Rt_scene, "Rt_scene").reshape(B, 16) if Rt_scene is not None else None
    if n_valid is not None:
        if not n_valid.is_cuda or n_valid.dtype != torch.int32 or n_valid.numel() != B:
            raise RuntimeError("fepe_b200: n_valid must be a CUDA int32 tensor of B elements")
        n_valid = n_valid.contiguous()
    with torch.cuda.device(E.device):
        out = torch.empty(L, B, _lib.RECOVER_OUT_FLOATS, dtype=torch.float32, device=E.device)
        mask = torch.empty(L, B, N, dtype=torch.uint8, device=E.device) if want_mask else None
        st = _lib.lib().fepe_recover_pose(E.data_ptr(), K.data_ptr(), matches.data_ptr(),
                                          n_valid.data_ptr() if n_valid is not None else None, L, B, N,
                                          float(distance_thresh), rt.data_ptr() if rt is not None else None,
                                          out.data_ptr(), mask.data_ptr() if mask is not None else None, _stream_ptr())
    _lib.check(st, "fepe_recover_pose")
    return out, mask


def gt_virt(K: torch.Tensor, Rt_scene: Optional[torch.Tensor], grid1: torch.Tensor, grid2: torch.Tensor,
            F_in: Optional[torch.Tensor] = None, want_normalized: bool = True):
    """Ground truth + virtual correspondences of a batch on the device (include/fepe_b200.h: fepe_gt_virt) -- per sample
    utils_F.E_F_from_Rt_np, utils_misc.get_virt_x1x2_np (cv2.correctMatches) and R_to_q_np in the reference's dataset
    (deepFEPE/datasets/kitti_odo_corr.py:290-302, :526-566).

    K [B,3,3]; Rt_scene [B,4,4] or None (then F_in [B,3,3] is required); grid1, grid2 [P,2] pixel grids.
    Returns (gt [B,32] | None, pts1_virt [B,P,3], pts2_virt [B,P,3], pts_virt_normalized [B,P,3] | None)."""
    K = _check_cuda_f32(K, "K")
    B = K.shape[0]
    K = K.reshape(B, 9)
    rt = _check_cuda_f32(Rt_scene, "Rt_scene").reshape(B, 16) if Rt_scene is not None else None
    fin = _check_cuda_f32(F_in, "F_in").reshape(B, 9) if F_in is not None else None
    if rt is None and fin is None:
        raise RuntimeError("fepe_b200: gt_virt needs Rt_scene or F_in")
    grid1 = _check_cuda_f32(grid1, "grid1")
    grid2 = _check_cuda_f32(grid2, "grid2")
    if grid1.dim() != 2 or grid1.shape[1] != 2 or grid1.shape != grid2.shape:
        raise RuntimeError("fepe_b200: grids must both be [P,2]")
    P = grid1.shape[0]
    dev = K.device
    with torch.cuda.device(dev):
        gt = torch.empty(B, _lib.GT_FLOATS, dtype=torch.float32, device=dev) if rt is not None else None
        p1 = torch.empty(B, P, 3, dtype=torch.float32, device=dev)
        p2 = torch.empty(B, P, 3, dtype=torch.float32, device=dev)
        pn = torch.empty(B, P, 3, dtype=torch.float32, device=dev) if want_normalized else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        st = _lib.lib().fepe_gt_virt(K.data_ptr(), ptr(rt), ptr(fin), grid1.data_ptr(), grid2.data_ptr(), B, P,
                                     ptr(gt), p1.data_ptr(), p2.data_ptr(), ptr(pn), _stream_ptr())
    _lib.check(st, "fepe_gt_virt")
    return gt, p1, p2, pn


def nn_match_two_way(desc1: torch.Tensor, desc2: torch.Tensor, nn_thresh: float,
                     n1: Optional[torch.Tensor] = None, n2: Optional[torch.Tensor] = None):
    """Mutual nearest-neighbour matching of L2-normalised descriptors for a whole batch (include/fepe_b200.h:
    fepe_nn_match) -- PointTracker.nn_match_two_way of the reference's call site train_good_utils.py:685-689.

    desc1 [B,N1,D], desc2 [B,N2,D] fp32 CUDA (one ROW per keypoint); n1, n2 [B] int32 valid counts or None.
    Returns idx1 [B,N1] int32, idx2 [B,N1] int32, score [B,N1] fp32, count [B] int32: the first count[b] columns of
    row b are the reference's matching_mask[0], [1], [2]."""
    desc1 = _check_cuda_f32(desc1, "desc1")
    desc2 = _check_cuda_f32(desc2, "desc2")
    if desc1.dim() != 3 or desc2.dim() != 3 or desc1.shape[0] != desc2.shape[0] or desc1.shape[2] != desc2.shape[2]:
        raise RuntimeError("fepe_b200: descriptors must be [B,N1,D] and [B,N2,D]")
    B, N1, D = desc1.shape
    N2 = desc2.shape[1]
    for name, t in (("n1", n1), ("n2", n2)):
        if t is not None and (not t.is_cuda or t.dtype != torch.int32 or t.numel() != B):
            raise RuntimeError(f"fepe_b200: {name} must be a CUDA int32 tensor of B elements")
    dev = desc1.device
    with torch.cuda.device(dev):
        ws = torch.empty(_lib.lib().fepe_nn_match_workspace_bytes(B, N1, N2) // 8, dtype=torch.int64, device=dev)
        idx1 = torch.empty(B, N1, dtype=torch.int32, device=dev)
        idx2 = torch.empty(B, N1, dtype=torch.int32, device=dev)
        score = torch.empty(B, N1, dtype=torch.float32, device=dev)
        count = torch.empty(B, dtype=torch.int32, device=dev)
        st = _lib.lib().fepe_nn_match(desc1.data_ptr(), desc2.data_ptr(),
                                      n1.contiguous().data_ptr() if n1 is not None else None,
                                      n2.contiguous().data_ptr() if n2 is not None else None,
                                      B, N1, N2, D, float(nn_thresh), ws.data_ptr(), idx1.data_ptr(), idx2.data_ptr(),
                                      score.data_ptr(), count.data_ptr(), _stream_ptr())
    _lib.check(st, "fepe_nn_match")
    return idx1, idx2, score, count


def pose_backward(F: torch.Tensor, K: torch.Tensor, affine, q_gt: torch.Tensor, t_gt: torch.Tensor,
                  virt1: Optional[torch.Tensor], virt2: Optional[torch.Tensor], clamp_at: float, pose_out: torch.Tensor,
                  g_q: Optional[torch.Tensor], g_t: Optional[torch.Tensor], g_loss: Optional[torch.Tensor],
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dL/dF [L,B,3,3] from the upstream gradients [L,B] of the q / t L2 errors and of the F-loss (any may be None):
    one launch of fepe_pose_bwd (include/fepe_b200.h).  `pose_out` is pose_forward's output for the same F."""
    F = _check_cuda_f32(F, "F")
    if F.dim() == 3:
        F = F.unsqueeze(0)
    L, B = F.shape[0], F.shape[1]
    prep = lambda g: _check_cuda_f32(g, "grad").reshape(L, B) if g is not None else None
    g_q, g_t, g_loss = prep(g_q), prep(g_t), prep(g_loss)
    Kc = _check_cuda_f32(K, "K").reshape(B, 9)
    qc = _check_cuda_f32(q_gt, "q_gt").reshape(B, 4)
    tc = _check_cuda_f32(t_gt, "t_gt").reshape(B, 3)
    has_v = virt1 is not None
    v1 = _check_cuda_f32(virt1, "virt1") if has_v else None
    v2 = _check_cuda_f32(virt2, "virt2") if has_v else None
    po = _check_cuda_f32(pose_out, "pose_out")
    with torch.cuda.device(F.device):
        dF = out if out is not None else torch.empty_like(F)
        ptr = lambda t: t.data_ptr() if t is not None else None
        st = _lib.lib().fepe_pose_bwd(F.data_ptr(), Kc.data_ptr(), L, B, affine[0], affine[1], affine[2], affine[3],
                                      qc.data_ptr(), tc.data_ptr(), ptr(v1), ptr(v2), v1.shape[1] if has_v else 0,
                                      float(clamp_at), po.data_ptr(), ptr(g_q), ptr(g_t), ptr(g_loss), dF.data_ptr(),
                                      _stream_ptr())
    _lib.check(st, "fepe_pose_bwd")
    return dF


class FitFunction(torch.autograd.Function):
    """Differentiable fused weighted 8-point fit.

    forward(matches [B,N,4], weights [B,N], ax, bx, ay, by, clamp_at) -> F [B,3,3], residual [B,N], epi [B,N]
    backward: d/d weights always; d/d matches (fepe_fit_bwd_coords) only when `matches` requires a gradient
    (if_learn_offsets / a trainable keypoint front-end in the reference)."""

    @staticmethod
    def forward(ctx, matches, weights, ax, bx, ay, by, clamp_at):
        aff = (float(ax), float(bx), float(ay), float(by))
        need = weights.requires_grad or matches.requires_grad
        F, res, epi, saved = fit_forward(matches, weights, aff, clamp_at, want_epi=True, want_saved=need)
        ctx.aff, ctx.clamp_at = aff, float(clamp_at)
        if need:
            ctx.save_for_backward(matches, weights, saved)
        ctx.mark_non_differentiable()
        return F, res, epi

    @staticmethod
    def backward(ctx, gF, gres, gepi):
        matches, weights, saved = ctx.saved_tensors
        B, N = matches.shape[0], matches.shape[1]
        gF = torch.zeros(B, 3, 3, device=matches.device) if gF is None else gF.contiguous()
        want_coords = ctx.needs_input_grad[0]
        g = fit_backward(matches, weights, saved, gF, gres.contiguous() if gres is not None else None,
                         gepi.contiguous() if gepi is not None else None, ctx.aff, ctx.clamp_at, want_coords)
        gw, gm = g if want_coords else (g, None)
        return gm, gw.reshape(weights.shape), None, None, None, None, None


class PoseLossFunction(torch.autograd.Function):
    """Differentiable pose / F-loss head on the device.

    forward(F [L,B,3,3], K, q_gt, t_gt, Rt_scene, virt1, virt2, ax, bx, ay, by, clamp_at)
        -> q_l2 [L,B], t_l2 [L,B], loss_F [L,B], out [L,B,32] (metrics, not differentiable)
    backward: dL/dF by fepe_pose_bwd (3x3 SVD adjoint, quaternion adjoint, epipolar-residual adjoint)."""

    @staticmethod
    def forward(ctx, F, K, q_gt, t_gt, Rt_scene, virt1, virt2, ax, bx, ay, by, clamp_at):
        aff = (float(ax), float(bx), float(ay), float(by))
        if F.dim() == 3:
            F = F.unsqueeze(0)
        F = F.contiguous()
        out = pose_forward(F, K, aff, q_gt, t_gt, Rt_scene, virt1, virt2, clamp_at)
        ctx.aff, ctx.clamp_at = aff, float(clamp_at)
        ctx.save_for_backward(F, K, q_gt, t_gt, virt1 if virt1 is not None else F.new_empty(0),
                              virt2 if virt2 is not None else F.new_empty(0), out)
        ctx.mark_non_differentiable(out)
        return out[..., 21].clone(), out[..., 22].clone(), out[..., 25].clone(), out

    @staticmethod
    def backward(ctx, g_q, g_t, g_loss, _g_out):
        F, K, q_gt, t_gt, virt1, virt2, out = ctx.saved_tensors
        has_v = virt1.numel() > 0
        dF = pose_backward(F, K, ctx.aff, q_gt, t_gt, virt1 if has_v else None, virt2 if has_v else None,
                           ctx.clamp_at, out, g_q, g_t, g_loss)
        return (dF,) + (None,) * 11


class LinearTC32(torch.autograd.Function):
    """y [M,Co] = x [M,K] @ W[Co,K]^T on tcgen05 at fp32 accuracy (split-fp16 operands, three MMAs per product, fp32
    accumulation; csrc/fepe_mlp32.cu), with the data gradient (the same GEMM with W^T) and the weight gradient
    (fepe_mlp32_wgrad, both operands read in place as MN-major tiles) on tensor cores as well.
    M % 128 == 0, K % 64 == 0, Co % 128 == 0 (use torch for thinner layers)."""

    @staticmethod
    def forward(ctx, x, W):
        from . import mlp32
        x = _check_cuda_f32(x, "x")
        W2 = _check_cuda_f32(W, "W")
        M, K = x.shape
        Co = W2.shape[0]
        if M % 128 or K % 64 or Co % 128 or W2.shape[1] != K:
            raise RuntimeError("fepe_b200.LinearTC32: need M % 128 == 0, K % 64 == 0, Co % 128 == 0")
        lib = _lib.lib()
        with torch.cuda.device(x.device):
            st = _stream_ptr()
            whi, wlo, wsc = mlp32.split_weight(lib, W2, st)
            y = torch.empty(M, Co, dtype=torch.float32, device=x.device)
            _lib.check(lib.fepe_mlp32_gemm(x.data_ptr(), None, 1.0, None, whi.data_ptr(), wlo.data_ptr(), wsc.data_ptr(),
                                           None, y.data_ptr(), None, 1, M, M, K, Co, st), "fepe_mlp32_gemm")
        ctx.save_for_backward(x, W2)
        return y

    @staticmethod
    def backward(ctx, gy):
        from . import mlp32
        x, W2 = ctx.saved_tensors
        M, K = x.shape
        Co = W2.shape[0]
        gy = _check_cuda_f32(gy, "gy")
        lib = _lib.lib()
        gx = gw = None
        with torch.cuda.device(x.device):
            st = _stream_ptr()
            amax = gy.abs().max().reshape(1).view(torch.int32)          # bit pattern of max |gy| (stays on the device)
            if ctx.needs_input_grad[0]:
                thi, tlo, tsc = mlp32.split_weight(lib, W2.t().contiguous(), st)
                gx = torch.empty(M, K, dtype=torch.float32, device=x.device)
                _lib.check(lib.fepe_mlp32_gemm(gy.data_ptr(), None, 1.0, amax.data_ptr(), thi.data_ptr(), tlo.data_ptr(),
                                               tsc.data_ptr(), None, gx.data_ptr(), None, 1, M, M, Co, K, st),
                           "fepe_mlp32_gemm(dgrad)")
            if ctx.needs_input_grad[1]:
                ident = torch.zeros(1, K, 2, dtype=torch.float32, device=x.device)     # (scale 1, shift 0) per channel
                ident[:, :, 0] = 1.0
                gw = torch.zeros(Co, K, dtype=torch.float32, device=x.device)
                _lib.check(lib.fepe_mlp32_wgrad(gy.data_ptr(), amax.data_ptr(), x.data_ptr(), ident.data_ptr(), 1.0,
                                                gw.data_ptr(), M, M, Co, K, st), "fepe_mlp32_wgrad")
        return gx, gw


def linear_tc32(x: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """x [M,K] @ W[Co,K]^T at fp32 accuracy on tensor cores (see LinearTC32)."""
    return LinearTC32.apply(x, W)

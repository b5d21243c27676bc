"""DeepFNet / Fit / NormalizeAndExpand_HW with the reference's module surface
(deepFEPE/models/DeepFNet.py) and the weighted 8-point path in hand-written CUDA.

Drop-in contract: same constructor arguments, same ``forward(data_batch) -> dict``
keys, same sub-module names (``input_weights``, ``update_weights``, ``norm_HW``, ``fit``) and hence the
same state_dict keys, so ``deepFEPE/utils/loader.py:modelLoader``, ``Train_model_pipeline`` and
``train_good_utils.get_all_loss_DeepF`` work unchanged on the returned dict.

What differs from the reference, deliberately:
  * ``if_cpu_svd`` is accepted and ignored -- there is no host round trip at all;
  * the sign of the null vector f is canonical (largest entry positive) instead of LAPACK's arbitrary
    one; F and the signed ``residual`` follow it, everything else is sign invariant;
  * gradients flow to the WEIGHTS (hence to both MLPs) and, when the coordinates require a gradient
    (``if_learn_offsets``, a trainable keypoint front-end), to the COORDINATES as well
    (``fepe_fit_bwd_coords``: Hartley transforms, row normalisation, eigenvector, epipolar distance);
  * the broken reference options ``if_des``, ``if_tri_depth`` are rejected.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .ErrorEstimators import ErrorEstimator
from .GoodCorresNet import GoodCorresNet


class NormalizeAndExpand_HW(nn.Module):
    """pixels -> [-1,1]^2 for both images (deepFEPE/models/DeepFNet.py:93-120)."""

    def __init__(self, image_size, is_cuda=True, is_test=False):
        super().__init__()
        self.H, self.W = image_size[0], image_size[1]
        self._T_cache = {}

    def affine(self):
        return ops.hw_affine((self.H, self.W))

    def _T(self, device, dtype):
        # cached per device: building it from a Python list is a blocking host-to-device copy, which would also make
        # the forward impossible to capture in a CUDA graph
        key = (device, dtype)
        T = self._T_cache.get(key)
        if T is None:
            T = torch.tensor([[2. / self.W, 0., -1.], [0., 2. / self.H, -1.], [0., 0., 1.]], device=device, dtype=dtype)
            self._T_cache[key] = T
        return T

    def normalize(self, pts):
        T = self._T(pts.device, pts.dtype).unsqueeze(0).expand(pts.size(0), -1, -1)
        ones = torch.ones(pts.size(0), pts.size(1), 1, device=pts.device, dtype=pts.dtype)
        return T @ torch.cat((pts, ones), 2).permute(0, 2, 1), T

    def forward(self, pts):
        pts1, T1 = self.normalize(pts[:, :, :2])
        pts2, T2 = self.normalize(pts[:, :, 2:])
        return pts1, pts2, T1, T2


class Fit(nn.Module):
    """Weighted 8-point fit (deepFEPE/models/DeepFNet.py:123-295): forward(pts1 [B,N,3], pts2 [B,N,3],
    weights [B,1,N]) -> (F [B,3,3], residual [B,N]).  One launch of the fused kernel."""

    def __init__(self, is_cuda=True, is_test=False, if_cpu_svd=False, normalize_SVD=True):
        super().__init__()
        if not normalize_SVD:
            raise NotImplementedError("normalize_SVD=False is never used by the reference (DeepFNet.py:353)")

    def forward(self, pts1, pts2, weights, if_print=False, matches_good_unique_num=None):
        # the kernel takes (x1,y1,x2,y2) rows; homogeneous inputs are assumed to have z = 1 as in every
        # call site of the reference (NormalizeAndExpand_HW keeps the third row [0,0,1])
        matches = torch.cat((pts1[:, :, :2], pts2[:, :, :2]), 2).contiguous()
        B, N = matches.shape[0], matches.shape[1]
        out, residual, _ = ops.FitFunction.apply(matches, weights.reshape(B, N), 1.0, 0.0, 1.0, 0.0, 0.5)
        return out, residual


class DeepFNet(nn.Module):
    def __init__(self, depth, image_size, if_quality, if_img_w=False, if_goodCorresArch=False, if_tri_depth=False,
                 if_learn_offsets=False, if_des=False, des_size=None, quality_size=0, is_cuda=True, is_test=False,
                 if_cpu_svd=False, **params):
        super().__init__()
        if if_des or if_tri_depth:
            raise NotImplementedError("if_des / if_tri_depth are broken in the reference itself "
                                      "(DeepFNet.py:484,508) and are not provided")
        if not if_quality:
            quality_size = 0
        self.if_quality = if_quality
        self.if_img_w = if_img_w
        self.if_goodCorresArch = if_goodCorresArch
        self.if_learn_offsets = bool(if_learn_offsets)
        self.image_size = image_size
        self.depth = depth
        if if_goodCorresArch:
            # the reference builds 4+Q / 6+Q channel nets here but feeds 7+Q channels (DeepFNet.py:337,487):
            # the update net is given the channel count it actually receives
            self.input_weights = GoodCorresNet(4 + quality_size, bn=False)
            self.update_weights = GoodCorresNet(4 + quality_size + 3, bn=False)
        else:
            self.input_weights = ErrorEstimator(4 + quality_size)
            self.update_weights = ErrorEstimator(4 + quality_size + 3)   # + weights, epi_res, residual
            if if_learn_offsets:      # DeepFNet.py:341-342
                self.update_offsets = ErrorEstimator(4 + quality_size + 3, output_size=4, if_bn=False)
        if is_test:
            self.input_weights.eval()
            self.update_weights.eval()
            if self.if_learn_offsets and hasattr(self, "update_offsets"):
                self.update_offsets.eval()
        self.norm_HW = NormalizeAndExpand_HW(self.image_size, is_cuda, is_test)
        self.fit = Fit(is_cuda, is_test, if_cpu_svd)

    def enable_tensor_core_mlp(self, inference: bool = True, training: bool = False):
        """Round-1 switch, kept for callers: the opt-in bf16 tcgen05 path (faster, outside the reference's fp32
        tolerance).  `set_mlp_path("tc32")` -- fp32-parity tensor cores -- is the default."""
        return self.set_mlp_path("bf16" if (inference or training) else "tc32")

    def set_mlp_path(self, path: str):
        """Select the arithmetic of every weight network: "tc32" (split-fp16 tcgen05 GEMMs at fp32 parity, the default),
        "bf16" (the faster bf16 tcgen05 path; opt-in, outside the reference's fp32 tolerance) or "torch" (PyTorch's fp32
        library kernels; A/B baseline only)."""
        for name in ("input_weights", "update_weights", "update_offsets"):
            net = getattr(self, name, None)
            if isinstance(net, ErrorEstimator):
                net.set_path(path)
        return self

    def get_input(self, data_batch, offsets=None, iter=None):
        pts = data_batch['matches_xy_ori']
        if offsets is not None:                       # DeepFNet.py:369-373
            pts = pts + offsets.permute(0, 2, 1)
        pts1, pts2, T1, T2 = self.norm_HW(pts)
        pts1 = pts1.permute(0, 2, 1)
        pts2 = pts2.permute(0, 2, 1)
        feats = [(pts1[:, :, :2] + 1) / 2, (pts2[:, :, :2] + 1) / 2]
        if self.if_quality:
            feats.append(data_batch['quality'])
        weight_in = torch.cat(feats, 2).permute(0, 2, 1)
        return weight_in, pts1, pts2, T1, T2

    @staticmethod
    def _softmax(net, logits):
        # the kernel paths compute the softmax over N in their last kernel
        sm = getattr(net, "last_softmax", None)
        return sm if sm is not None else F.softmax(logits, dim=2)

    @staticmethod
    def _net(net, matches, aff, extras, B, N):
        """Evaluate a weight network on [the 4 normalised coordinates of `matches`] + `extras` (channel groups in the
        order of the reference's torch.cat).  ErrorEstimator reads them in place (no cat / permute on its kernel paths)."""
        if isinstance(net, ErrorEstimator):
            return net.forward_parts(matches, aff, extras, B, N)
        return net(ErrorEstimator._features(matches, aff, extras))

    def forward(self, data_batch):
        matches = data_batch['matches_xy_ori']
        if not matches.is_cuda:
            raise RuntimeError("fepe_b200.DeepFNet needs CUDA tensors: there is no CPU path")
        matches = matches.float().contiguous()
        B, N = matches.shape[0], matches.shape[1]
        aff = self.norm_HW.affine()
        _, pts1, pts2, T1, T2 = self.get_input(data_batch)        # returned to the caller (outs['pts1'], ['T1'], ...)
        base = [data_batch['quality'].float()] if self.if_quality else []      # DeepFNet.py:385-389

        logits = self._net(self.input_weights, matches, aff, base, B, N)
        weights_pts = self._softmax(self.input_weights, logits)
        weights_prod = weights_pts * data_batch['weights_im'] if self.if_img_w else weights_pts
        _ = data_batch.get('matches_good_unique_nums'), data_batch.get('t_scene_scale')   # read, unused (:449,:453)

        out_layers, epi_res_layers, residual_layers = [], [], []
        weights_layers, logits_layers = [weights_prod], [logits]
        for _it in range(self.depth - 1):
            # fused: Fit.forward (:466) + compute_epi_residual(pts1, pts2, out) (:479), one kernel
            out, residual, epi = ops.FitFunction.apply(matches, weights_prod.reshape(B, N), *aff, 0.5)
            out_layers.append(out)
            residual_layers.append(residual)
            epi_res_layers.append(epi.unsqueeze(1))
            # net_in = cat(pts_normalized_in, weights_prod, epi_res, residual) (:487), read in place by the first layer
            extras = base + [weights_prod.reshape(B, N), epi, residual]
            if self.if_learn_offsets:                 # DeepFNet.py:489-505: offsets replace (not accumulate)
                offsets_accu = self._net(self.update_offsets, matches, aff, extras, B, N)
                _, pts1, pts2, T1, T2 = self.get_input(data_batch, offsets_accu, _it)
                matches = (data_batch['matches_xy_ori'].float() + offsets_accu.permute(0, 2, 1)).contiguous()
            logits = self._net(self.update_weights, matches, aff, extras, B, N)
            weights_pts = self._softmax(self.update_weights, logits)
            weights_prod = weights_pts * data_batch['weights_im'] if self.if_img_w else weights_pts
            weights_layers.append(weights_prod)
            logits_layers.append(logits)

        out, residual, _ = ops.FitFunction.apply(matches, weights_prod.reshape(B, N), *aff, 0.5)
        residual_layers.append(residual)
        out_layers.append(out)
        preds = {
            "logits": logits.squeeze(1), 'logits_layers': logits_layers, 'F_est': out,
            'epi_res_layers': epi_res_layers, 'T1': T1, 'T2': T2, 'out_layers': out_layers,
            'pts1': pts1, 'pts2': pts2, 'weights': weights_prod, 'residual_layers': residual_layers,
            'weights_layers': weights_layers,
        }
        if self.if_learn_offsets and self.depth > 1:
            preds['offsets'] = offsets_accu           # DeepFNet.py:549-550
        return preds

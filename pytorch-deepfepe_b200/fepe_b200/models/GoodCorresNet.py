"""GoodCorresNet -- PointNet part-segmentation style weight network
(deepFEPE/models/GoodCorresNet.py:35-163, marked "deprecated" there).

The reference file cannot be instantiated as shipped: Stem / SharedMLP / Conv1d / set_bn come from the
un-vendored `haosulab/shaper` package whose import is commented out (:14-21).  The blocks are
re-created here from the constructor's channel specification (:45-53) and the data flow of forward
(:95-160): stem 64/128/128 -> local MLP 512/2048 -> max over N -> [stem feats + local feats + global]
(4928 channels) -> seg MLP 256/256 (dropout) -> 128 -> 1 logit per correspondence.
PARITY UNPINNED: there is no reference output to compare with; the tests pin the module against its own
fp64 evaluation and a committed self-golden.

Arithmetic: every 1x1 convolution that is a genuine dense GEMM (K % 64 == 0, Co % 128 == 0: all but the first and the
last layer, 99.9 % of the flops) runs on tcgen05 at fp32 accuracy through ops.linear_tc32 -- forward, data gradient
and weight gradient -- in a points-major [B * Npad, C] layout (Npad = N rounded up to 128; padded rows are masked out
of the statistics and of the max-pool).  InstanceNorm / ReLU / the global max-pool / the concatenation are PyTorch
elementwise glue.  `use_kernels = False` evaluates the same module with PyTorch's fp32 matmul (A/B baseline).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _Conv1dBlock(nn.Module):
    def __init__(self, cin, cout, with_instance_norm=True, relu=True, dropout=0.0):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, 1, bias=not with_instance_norm)
        self.norm = nn.InstanceNorm1d(cout, affine=True) if with_instance_norm else None
        self.relu = relu
        self.drop = nn.Dropout(dropout) if dropout > 0 else None

    def forward(self, x, n_valid, use_kernels):
        """x [B, Npad, Cin] points-major (rows >= n_valid are padding) -> [B, Npad, Cout]."""
        B, Npad, cin = x.shape
        W = self.conv.weight.reshape(self.conv.out_channels, cin)
        if use_kernels and x.is_cuda and cin % 64 == 0 and W.shape[0] % 128 == 0:
            from .. import ops
            y = ops.linear_tc32(x.reshape(B * Npad, cin), W).reshape(B, Npad, -1)
        else:
            y = x @ W.t()
        if self.conv.bias is not None:
            y = y + self.conv.bias
        if self.norm is not None:                      # InstanceNorm1d over the n_valid real points of each pair
            v = y[:, :n_valid]
            mean = v.mean(1, keepdim=True)
            var = v.var(1, unbiased=False, keepdim=True)
            y = (y - mean) * torch.rsqrt(var + self.norm.eps) * self.norm.weight + self.norm.bias
        if self.relu:
            y = F.relu(y)
        if self.drop is not None:
            y = self.drop(y)
        if n_valid < Npad:
            y = torch.cat((y[:, :n_valid], y.new_zeros(B, Npad - n_valid, y.shape[2])), 1)
        return y


class _SharedMLP(nn.ModuleList):
    def __init__(self, cin, channels, dropout_prob=0.0, with_instance_norm=True):
        super().__init__()
        for c in channels:
            self.append(_Conv1dBlock(cin, c, with_instance_norm, True, dropout_prob))
            cin = c
        self.out_channels = cin


class GoodCorresNet(nn.Module):
    def __init__(self, in_channels, num_classes=1, stem_channels=(64, 128, 128), local_channels=(512, 2048),
                 seg_channels=(256, 256, 128), dropout_prob=0.2, with_transform=False, bn=False):
        super().__init__()
        self.in_channels = in_channels
        self.stem = _SharedMLP(in_channels, stem_channels)
        self.mlp_local = _SharedMLP(stem_channels[-1], local_channels)
        cat_ch = sum(stem_channels) + sum(local_channels) + local_channels[-1]      # 4928
        self.mlp_seg = _SharedMLP(cat_ch, seg_channels[:-1], dropout_prob=dropout_prob)
        self.conv_seg = _Conv1dBlock(seg_channels[-2], seg_channels[-1])
        self.seg_logit = nn.Conv1d(seg_channels[-1], num_classes, 1, bias=True)
        self.use_kernels = True

    def forward(self, x):
        """x [B, Cin, N] -> logits [B, num_classes, N] (the reference's interface, GoodCorresNet.py:95-160)."""
        if not x.is_cuda:
            raise RuntimeError("fepe_b200.GoodCorresNet needs CUDA tensors: there is no CPU path")
        return self._forward(x, self.use_kernels)

    def _forward(self, x, use_kernels):
        B, _, N = x.shape
        Npad = (N + 127) // 128 * 128
        h = x.permute(0, 2, 1)
        if Npad > N:
            h = torch.cat((h, h.new_zeros(B, Npad - N, h.shape[2])), 1)
        feats = []
        for m in self.stem:
            h = m(h, N, use_kernels)
            feats.append(h)
        for m in self.mlp_local:
            h = m(h, N, use_kernels)
            feats.append(h)
        g = h[:, :N].max(1, keepdim=True).values                     # global feature: max over the real points
        feats.append(g.expand(-1, Npad, -1))
        h = torch.cat(feats, 2)                                        # [B, Npad, 4928]
        for m in self.mlp_seg:
            h = m(h, N, use_kernels)
        h = self.conv_seg(h, N, use_kernels)
        logits = h[:, :N] @ self.seg_logit.weight.reshape(self.seg_logit.out_channels, -1).t() + self.seg_logit.bias
        return logits.permute(0, 2, 1)

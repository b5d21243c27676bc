"""GoodCorresNet -- PointNet part-segmentation style weight network
(deepFEPE/models/GoodCorresNet.py:35-163, marked "deprecated" there).

The reference file cannot be instantiated as shipped: Stem / SharedMLP / Conv1d / set_bn come from the
un-vendored `haosulab/shaper` package whose import is commented out (:14-21).  The blocks are
re-created here from the constructor's channel specification (:45-53) and the data flow of forward
(:95-160): stem 64/128/128 -> local MLP 512/2048 -> max over N -> [stem feats + local feats + global]
(4928 channels) -> seg MLP 256/256 (dropout) -> 128 -> 1 logit per correspondence.
PARITY UNPINNED: there is no reference output to compare with.
"""
import torch
import torch.nn as nn


class _Conv1dBlock(nn.Module):
    def __init__(self, cin, cout, with_instance_norm=True, relu=True, dropout=0.0):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, 1, bias=not with_instance_norm)
        self.norm = nn.InstanceNorm1d(cout, affine=True) if with_instance_norm else None
        self.relu = nn.ReLU(inplace=True) if relu else None
        self.drop = nn.Dropout(dropout) if dropout > 0 else None

    def forward(self, x):
        x = self.conv(x)
        if self.norm is not None:
            x = self.norm(x)
        if self.relu is not None:
            x = self.relu(x)
        if self.drop is not None:
            x = self.drop(x)
        return x


class _SharedMLP(nn.ModuleList):
    def __init__(self, cin, channels, dropout_prob=0.0, with_instance_norm=True):
        super().__init__()
        for c in channels:
            self.append(_Conv1dBlock(cin, c, with_instance_norm, True, dropout_prob))
            cin = c
        self.out_channels = cin

    def forward(self, x):
        for m in self:
            x = m(x)
        return x


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

    def forward(self, x):
        N = x.shape[2]
        feats = []
        for m in self.stem:
            x = m(x)
            feats.append(x)
        for m in self.mlp_local:
            x = m(x)
            feats.append(x)
        g, _ = torch.max(x, 2, keepdim=True)
        feats.append(g.expand(-1, -1, N))
        x = torch.cat(feats, 1)
        x = self.mlp_seg(x)
        x = self.conv_seg(x)
        return self.seg_logit(x)

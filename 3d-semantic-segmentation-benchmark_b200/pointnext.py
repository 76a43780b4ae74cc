"""PointNeXt-style segmentation net -- caller of the hot path, same layers and state_dict keys as
/root/reference/models/PointNeXt/PointNeXt.py:17-147 (`version` is accepted and unused there too)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .common import FeaturePropagation, InvResMLP, SetAbstraction, UnitPointNet, head_dropout_p


class PointNeXt(nn.Module):
    def __init__(self, part_classes: int, version: str = 'b'):
        super().__init__()
        self.num_classes = part_classes
        self.mlp = UnitPointNet(9, [32])
        self.sa1 = SetAbstraction(1024, 0.1, 32 + 3, [32, 32, 64], grouping_norm=True)
        self.irmlp1 = InvResMLP(0.1, 64 + 3, 64, 32)
        self.sa2 = SetAbstraction(256, 0.2, 64 + 3, [64, 64, 128], grouping_norm=True)
        self.irmlp2 = InvResMLP(0.1, 128 + 3, 128, 32)
        self.irmlp2_1 = InvResMLP(0.2, 128 + 3, 128, 32)
        self.sa3 = SetAbstraction(64, 0.4, 128 + 3, [128, 128, 256], grouping_norm=True)
        self.irmlp3 = InvResMLP(0.4, 256 + 3, 256, 32)
        self.sa4 = SetAbstraction(16, 0.8, 256 + 3, [256, 256, 512], grouping_norm=True)
        self.irmlp4 = InvResMLP(0.8, 512 + 3, 512, 16)
        self.fp4 = FeaturePropagation(512 + 256, [256, 256])
        self.fp3 = FeaturePropagation(256 + 128, [256, 256])
        self.fp2 = FeaturePropagation(256 + 64, [256, 128])
        self.fp1 = FeaturePropagation(128 + 32, [128, 128, 128, 128])
        self.drop = nn.Dropout(0.5)
        self.conv = nn.Conv1d(128, part_classes, 1)

    def forward(self, x: torch.Tensor, lengths=None) -> torch.Tensor:
        """x (B,N,9) -> raw logits (B,N,part_classes).  lengths (B,): length-aware evaluation of a zero-padded batch, as
        PointNetpp.forward (only the input level has padding rows: every deeper level is made of real centroids)."""
        xt = x.permute(0, 2, 1)
        coords_0 = xt[:, :3, :].permute(0, 2, 1)
        features_0 = self.mlp(xt).permute(0, 2, 1)
        coords_1, features_1 = self.sa1(coords_0, features_0, lengths=lengths)
        coords_1, features_1 = self.irmlp1(coords_1, coords_1, features_1)
        coords_2, features_2 = self.sa2(coords_1, features_1)
        coords_2, features_2 = self.irmlp2(coords_2, coords_2, features_2)
        coords_2, features_2 = self.irmlp2_1(coords_2, coords_2, features_2)
        coords_3, features_3 = self.sa3(coords_2, features_2)
        coords_3, features_3 = self.irmlp3(coords_3, coords_3, features_3)
        coords_4, features_4 = self.sa4(coords_3, features_3)
        coords_4, features_4 = self.irmlp4(coords_4, coords_4, features_4)
        features_3 = self.fp4(coords_3, coords_4, features_3, features_4)
        features_2 = self.fp3(coords_2, coords_3, features_2, features_3)
        features_1 = self.fp2(coords_1, coords_2, features_1, features_2)
        p = head_dropout_p(self.drop)                   # nn.Dropout behind fp1 (PointNeXt.py:134): folded into its last fused layer
        features_0 = self.fp1(coords_0, coords_1, features_0, features_1, lengths=lengths, _dropout=p)
        x = features_0 if p > 0.0 else self.drop(features_0)     # (B,N,128): the head 1x1 conv is a GEMM over the rows
        return ops.linear_rows(x, self.conv.weight.squeeze(-1), self.conv.bias)

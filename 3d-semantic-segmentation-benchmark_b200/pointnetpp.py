"""PointNet++ SSG semantic segmentation -- caller of the hot path, same layers and state_dict keys as
/root/reference/models/PointNetpp/PointNetpp.py:6-48."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .common import FeaturePropagation, SetAbstraction


class PointNetpp(nn.Module):
    def __init__(self, part_classes: int):
        super().__init__()
        self.sa1 = SetAbstraction(1024, 0.1, 9, [32, 32, 64])
        self.sa2 = SetAbstraction(256, 0.2, 64 + 3, [64, 64, 128])
        self.sa3 = SetAbstraction(64, 0.4, 128 + 3, [128, 128, 256])
        self.sa4 = SetAbstraction(16, 0.8, 256 + 3, [256, 256, 512])
        self.fp4 = FeaturePropagation(512 + 256, [256, 256])
        self.fp3 = FeaturePropagation(256 + 128, [256, 256])
        self.fp2 = FeaturePropagation(256 + 64, [256, 128])
        self.fp1 = FeaturePropagation(128, [128, 128, 128, 128])
        self.drop = nn.Dropout(0.5)
        self.conv = nn.Conv1d(128, part_classes, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x (B,N,9): xyz, rgb, block-centred xyz -> raw logits (B,N,part_classes)."""
        coords_0, features_0 = x[:, :, :3], x[:, :, 3:]
        sas = (self.sa1, self.sa2, self.sa3, self.sa4)
        # every index of the network depends on coordinates only: FPS / ball query of the deeper levels and the decoder's
        # 3-NN tables run on a side stream while the feature path of the shallower levels computes (same picks, same
        # generator consumption as calling sample() level by level)
        geo = ops.PyramidGeometry(coords_0, [(m.C, m.radius, m.K) for m in sas], [m.fps_start for m in sas])
        coords_1, features_1 = self.sa1(geo.coords[0], features_0, _geom=geo.level(0))
        coords_2, features_2 = self.sa2(coords_1, features_1, _geom=geo.level(1))
        coords_3, features_3 = self.sa3(coords_2, features_2, _geom=geo.level(2))
        coords_4, features_4 = self.sa4(coords_3, features_3, _geom=geo.level(3))
        features_3 = self.fp4(coords_3, coords_4, features_3, features_4, _geom=geo.three_nn(3))
        features_2 = self.fp3(coords_2, coords_3, features_2, features_3, _geom=geo.three_nn(2))
        features_1 = self.fp2(coords_1, coords_2, features_1, features_2, _geom=geo.three_nn(1))
        features_0 = self.fp1(coords_0, coords_1, None, features_1, _geom=geo.three_nn(0))
        x = self.drop(features_0)                       # (B,N,128): the head 1x1 conv is a GEMM over the rows
        return ops.linear_rows(x, self.conv.weight.squeeze(-1), self.conv.bias)

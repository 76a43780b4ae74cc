"""PointNet++ SSG semantic segmentation -- caller of the hot path, same layers and state_dict keys as
/root/reference/models/PointNetpp/PointNetpp.py:6-48."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .common import FeaturePropagation, SetAbstraction, SetAbstractionMSG, head_dropout_p


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

    def prepare_geometry(self, x: torch.Tensor, stream=None):
        """Every index of the network for the batch x (B,N,9) -- FPS picks, ball-query tables, the decoder's 3-NN tables and
        the CSR inverses their backward kernels need -- as a flat list of tensors (ops.PyramidGeometry.export).  It depends
        on the coordinates only, so a training loop can compute it for the NEXT batch on the side stream while the current
        step runs (train.GraphedTrainStep(geometry_fn=...)) and hand it to forward(x, geometry=...).  stream: run the
        kernels there (the caller waits on it before using the tensors)."""
        sas = (self.sa1, self.sa2, self.sa3, self.sa4)
        geo = ops.PyramidGeometry(x[:, :, :3], [(m.C, m.radius, m.K) for m in sas], [m.fps_start for m in sas], inline=True,
                                  stream=stream)
        out = geo.export()
        out.append(geo)                                  # keeps the workspaces alive until the caller has copied the tensors
        return out

    def forward(self, x: torch.Tensor, geometry=None, lengths=None) -> torch.Tensor:
        """x (B,N,9): xyz, rgb, block-centred xyz -> raw logits (B,N,part_classes).  geometry: prepare_geometry(x)'s tensors
        (without the trailing keep-alive object) when they were computed ahead of time.
        lengths (B,): length-aware evaluation of a zero-padded batch (the loader's third output,
        data_processing/block_datasets.py:27): the padding rows take no part in FPS / grouping / 3-NN, and in
        eval mode the logits of the real rows are those of the reference on each cloud passed alone.  None = the
        reference's behaviour (the padding participates, Training/training.py:112)."""
        coords_0, features_0 = x[:, :, :3], x[:, :, 3:]
        sas = (self.sa1, self.sa2, self.sa3, self.sa4)
        # every index of the network depends on coordinates only: FPS / ball query of the deeper levels and the decoder's
        # 3-NN tables run on a side stream while the feature path of the shallower levels computes (same picks, same
        # generator consumption as calling sample() level by level)
        if geometry is not None:
            geo = ops.PyramidGeometry.from_export(coords_0, 4, geometry)
        else:
            geo = ops.PyramidGeometry(coords_0, [(m.C, m.radius, m.K) for m in sas], [m.fps_start for m in sas], lengths=lengths)
        coords_1, features_1 = self.sa1(geo.coords[0], features_0, _geom=geo.level(0))
        coords_2, features_2 = self.sa2(coords_1, features_1, _geom=geo.level(1))
        coords_3, features_3 = self.sa3(coords_2, features_2, _geom=geo.level(2))
        coords_4, features_4 = self.sa4(coords_3, features_3, _geom=geo.level(3))
        features_3 = self.fp4(coords_3, coords_4, features_3, features_4, _geom=geo.three_nn(3))
        features_2 = self.fp3(coords_2, coords_3, features_2, features_3, _geom=geo.three_nn(2))
        features_1 = self.fp2(coords_1, coords_2, features_1, features_2, _geom=geo.three_nn(1))
        p = head_dropout_p(self.drop)                   # nn.Dropout behind fp1 (PointNetpp.py:42): folded into its last fused layer
        features_0 = self.fp1(coords_0, coords_1, None, features_1, _geom=geo.three_nn(0), _dropout=p)
        x = features_0 if p > 0.0 else self.drop(features_0)     # (B,N,128): the head 1x1 conv is a GEMM over the rows
        return ops.linear_rows(x, self.conv.weight.squeeze(-1), self.conv.bias)


class PointNetppMSG(nn.Module):
    """PointNet++ MSG semantic segmentation (BASELINE configs[2]: multi-radius ball query, batch 32 x 4096 points).

    The reference ships only the SSG network (models/PointNetpp/PointNetpp.py:6-48); this is the same
    encoder / decoder skeleton with every set abstraction replaced by its multi-scale form -- two radii per level
    (r/2 with K=16 and r with K=32 for the SSG radii r = 0.1 / 0.2 / 0.4 / 0.8, as in the original PointNet++ MSG
    segmentation network; each scale's MLP is the SSG MLP of that level, the small scale of level 1 half as wide) --
    built from the reference's own blocks (MiniPointNet, FeaturePropagation)."""

    def __init__(self, part_classes: int):
        super().__init__()
        self.sa1 = SetAbstractionMSG(1024, [0.05, 0.1], 9, [[16, 16, 32], [32, 32, 64]], [16, 32])
        self.sa2 = SetAbstractionMSG(256, [0.1, 0.2], 96 + 3, [[64, 64, 128], [64, 64, 128]], [16, 32])
        self.sa3 = SetAbstractionMSG(64, [0.2, 0.4], 256 + 3, [[128, 128, 256], [128, 128, 256]], [16, 32])
        self.sa4 = SetAbstractionMSG(16, [0.4, 0.8], 512 + 3, [[256, 256, 512], [256, 256, 512]], [16, 32])
        self.fp4 = FeaturePropagation(1024 + 512, [256, 256])
        self.fp3 = FeaturePropagation(256 + 256, [256, 256])
        self.fp2 = FeaturePropagation(256 + 96, [256, 128])
        self.fp1 = FeaturePropagation(128, [128, 128, 128, 128])
        self.drop = nn.Dropout(0.5)
        self.conv = nn.Conv1d(128, part_classes, 1)

    def forward(self, x: torch.Tensor, lengths=None) -> torch.Tensor:
        """x (B,N,9): xyz, rgb, block-centred xyz -> raw logits (B,N,part_classes).  lengths: as PointNetpp.forward."""
        coords_0, features_0 = x[:, :, :3], x[:, :, 3:]
        sas = (self.sa1, self.sa2, self.sa3, self.sa4)
        # as in PointNetpp.forward: the deeper FPS levels, their multi-radius ball queries and the decoder's 3-NN tables run
        # on the side stream while the feature path of the shallower levels computes
        geo = ops.PyramidGeometry(coords_0, [(m.C, m.radii, m.Ks) for m in sas], [m.fps_start for m in sas], lengths=lengths)
        coords_1, features_1 = self.sa1(geo.coords[0], features_0, _geom=geo.level(0))
        coords_2, features_2 = self.sa2(coords_1, features_1, _geom=geo.level(1))
        coords_3, features_3 = self.sa3(coords_2, features_2, _geom=geo.level(2))
        coords_4, features_4 = self.sa4(coords_3, features_3, _geom=geo.level(3))
        features_3 = self.fp4(coords_3, coords_4, features_3, features_4, _geom=geo.three_nn(3))
        features_2 = self.fp3(coords_2, coords_3, features_2, features_3, _geom=geo.three_nn(2))
        features_1 = self.fp2(coords_1, coords_2, features_1, features_2, _geom=geo.three_nn(1))
        p = head_dropout_p(self.drop)
        features_0 = self.fp1(coords_0, coords_1, None, features_1, _geom=geo.three_nn(0), _dropout=p)
        x = features_0 if p > 0.0 else self.drop(features_0)
        return ops.linear_rows(x, self.conv.weight.squeeze(-1), self.conv.bias)

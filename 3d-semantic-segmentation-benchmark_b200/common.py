"""Host-side mirror of the reference's models/utils/common.py on top of libpcnbr.

Same names, argument meaning, return layouts, parameter names (state_dict keys) and error behaviour
as /root/reference/models/utils/common.py, so PointNetpp / PointNeXt and train.py use it as a
drop-in; every neighbourhood op runs in a hand-written sm_100a kernel.  CUDA tensors only.

Selection follows the canonical tie rule (lowest index wins), which is what a stable sort of the
reference's distance rows yields; torch.topk's own tie order is unspecified.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

__all__ = ["sample", "group", "reduce", "interpolate", "MiniPointNet", "UnitPointNet", "SetAbstraction",
           "SetAbstractionMSG", "FeaturePropagation", "InvResMLP"]


def sample(coords: torch.Tensor, C: int, start_idx: torch.Tensor | None = None, lengths=None) -> torch.Tensor:
    """Farthest point sampling -> coordinates (B,C,3) of the picks   [reference common.py:6-34].

    The first pick is drawn like the reference (one torch.randint on coords.device) unless
    `start_idx` (B,) is given.  lengths (B,), here and below: the length-aware form for the zero-padded batches of the
    evaluation loader (data_processing/block_datasets.py:19-25) -- the padding rows take no part, the
    result for the real rows is the reference's on the cloud passed alone.  None = the reference's behaviour (padding
    participates)."""
    return ops.farthest_point_sample(coords, C, start_idx, return_coords=True, lengths=lengths)[1]


def group(centroid_coords: torch.Tensor, coords: torch.Tensor, features: torch.Tensor, r: float, K: int,
          normalize: bool = False, lengths=None) -> torch.Tensor:
    """Ball query + gather + centre-subtract (+ /r) + concat -> (B,C,K,3+D)   [common.py:37-71]."""
    nbr = ops.NeighborIndex(ops.query_ball_point(r, K, coords, centroid_coords, lengths=lengths), coords.shape[1])
    return ops.group_points(coords, features, centroid_coords, nbr, r if normalize else None)


def _group_rows(centroid_coords, coords, features, r, K, normalize, lengths=None):
    """group() for the modules below: the same tensor with its rows zero-padded to a multiple of 4 floats, the pitch
    the tensor-core GEMM of the first 1x1 convolution reads in place (MiniPointNet.forward_rows)."""
    nbr = ops.NeighborIndex(ops.query_ball_point(r, K, coords, centroid_coords, lengths=lengths), coords.shape[1])
    return ops.group_points(coords, features, centroid_coords, nbr, r if normalize else None, pad4=True)


def reduce(x: torch.Tensor, type: str) -> torch.Tensor:
    """Pooling over the K axis of (B,C,K,D')   [common.py:74-91].

    'avg' reproduces the reference literally, including its `[0]` (batch element 0 of the mean,
    common.py:89); it is unused by the models and runs as plain torch ops."""
    if type == 'max':
        return ops.max_pool_neighbors(x, 2)
    if type == 'avg':
        return torch.mean(x, dim=2)[0]
    raise ValueError(f"'{type}' pooling not supported; use 'max' or 'avg'.")


def head_dropout_p(drop: nn.Dropout) -> float:
    """Probability to fold into the preceding fused layer (training mode, 0 < p < 1), else 0: the caller applies the module."""
    return float(drop.p) if (drop.training and 0.0 < drop.p < 1.0) else 0.0


def interpolate(points: torch.Tensor, coords_1: torch.Tensor, coords_2: torch.Tensor, k: int = 3, lengths=None) -> torch.Tensor:
    """k-NN inverse-squared-distance interpolation (B,M,D) -> (B,N,D)   [common.py:94-122].  lengths: real rows of coords_1."""
    idx, d2 = ops.knn_points(coords_1, coords_2, k, query_lengths=lengths)
    return ops.three_interpolate(points, ops.NeighborIndex(idx, coords_2.shape[1]), d2)


class MiniPointNet(nn.Module):
    """[Conv2d 1x1 -> BatchNorm2d -> ReLU] x len(mlps)   [common.py:125-150]; library convolutions."""

    def __init__(self, in_channels: int, mlps: list[int]):
        super().__init__()
        self.conv = nn.ModuleList()
        self.batch = nn.ModuleList()
        width = in_channels
        for m in mlps:
            self.conv.append(nn.Conv2d(width, m, (1, 1)))
            self.batch.append(nn.BatchNorm2d(m))
            width = m

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x (B,Cin,C,K).  When x is the channels-last view the set-abstraction modules hand in (memory
        (B,C,K,Cin)), each 1x1 convolution is one cuBLAS SGEMM over the (B*C*K, Cin) rows and BatchNorm2d runs as
        a batch norm over the same rows (identical statistics); otherwise the plain cuDNN path is used."""
        B, _, C, K = x.shape
        rows = x.permute(0, 2, 3, 1)
        if not (x.is_cuda and rows.is_contiguous()):
            ops.note_fallback("MiniPointNet.forward: input is not a channels-last view (cuDNN convolutions)")
            for conv, bn in zip(self.conv, self.batch):
                x = F.relu(bn(conv(x)))
            return x
        return self.forward_rows(rows).permute(0, 3, 1, 2)

    def forward_rows(self, rows: torch.Tensor, pool_max: bool = False) -> torch.Tensor:
        """rows (B,C,K,Cin') point-major, Cin' = Cin or Cin padded with zero columns (ops.group_points(pad4=True))
        -> (B,C,K,Cout), or (B,C,Cout) = its max over K when pool_max (the last layer is then fused with the pooling).
        The first weight is zero-padded to match, so the result and all gradients are those of the unpadded layer."""
        h = rows
        last = len(self.conv) - 1
        for i, (conv, bn) in enumerate(zip(self.conv, self.batch)):
            w = conv.weight.view(conv.out_channels, conv.in_channels)
            if i == 0 and h.shape[-1] != conv.in_channels:
                w = F.pad(w, (0, h.shape[-1] - conv.in_channels))
            if pool_max and i == last:
                return ops.linear_bn_act_maxpool_rows(h, w, conv.bias, bn, 0.0)   # conv + BatchNorm2d + ReLU + max over K
            h = ops.linear_bn_act_rows(h, w, conv.bias, bn, 0.0)            # conv 1x1 + BatchNorm2d + ReLU (B,C,K,Cout)
        return h


def _batch_norm_rows(bn: nn.modules.batchnorm._BatchNorm, rows: torch.Tensor) -> torch.Tensor:
    """nn.BatchNorm{1,2}d applied to a (rows, channels) matrix: same statistics, running-stat update and
    num_batches_tracked bookkeeping as the module's own forward."""
    ops.note_fallback("F.batch_norm on rows")
    training = bn.training or bn.running_mean is None
    momentum = bn.momentum
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    return F.batch_norm(rows, bn.running_mean if (not training or bn.track_running_stats) else None,
                        bn.running_var if (not training or bn.track_running_stats) else None,
                        bn.weight, bn.bias, training, 0.0 if momentum is None else momentum, bn.eps)


class UnitPointNet(nn.Module):
    """[Conv1d -> BatchNorm1d -> ReLU] x len(mlps)   [common.py:153-178]."""

    def __init__(self, in_channels: int, mlps: list[int]):
        super().__init__()
        self.conv = nn.ModuleList()
        self.batch = nn.ModuleList()
        width = in_channels
        for m in mlps:
            self.conv.append(nn.Conv1d(width, m, 1))
            self.batch.append(nn.BatchNorm1d(m))
            width = m

    def forward(self, x: torch.Tensor, _dropout: float = 0.0) -> torch.Tensor:
        """x (B,Cin,N).  When x is the transposed view of point-major memory (B,N,Cin) -- what the feature
        propagation / InvResMLP modules hand in -- the 1x1 convolutions run as tensor-core GEMMs over the (B*N, Cin)
        rows; otherwise the plain cuDNN path is used.  _dropout (internal): probability of an nn.Dropout that FOLLOWS the
        module in training mode (the segmentation heads, PointNetpp.py:42, PointNeXt.py:134): folded into the last layer's fused
        BatchNorm + ReLU kernels (mask from a per-pass device seed, recomputed by the backward) instead of two more passes."""
        B, _, N = x.shape
        rows = x.permute(0, 2, 1)
        if not (x.is_cuda and rows.is_contiguous()):
            ops.note_fallback("UnitPointNet.forward: input is not a transposed point-major view (cuDNN convolutions)")
            for conv, bn in zip(self.conv, self.batch):
                x = F.relu(bn(conv(x)))
            return F.dropout(x, _dropout, True) if _dropout > 0.0 else x
        h = rows
        last = len(self.conv) - 1
        for i, (conv, bn) in enumerate(zip(self.conv, self.batch)):
            h = ops.linear_bn_act_rows(h, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias, bn, 0.0,
                                       _dropout if i == last else 0.0)
        return h.permute(0, 2, 1)


def _unit_forward_rows_cat(self, rows1: torch.Tensor, rows2: torch.Tensor, _dropout: float = 0.0) -> torch.Tensor:
    """UnitPointNet on the channel concatenation [rows1 | rows2] of two point-major (B,N,*) tensors -> (B,N,Cout): the
    skip connection of FeaturePropagation (torch.cat, common.py:234-237) is read by the first GEMM from the two tensors in
    place instead of being materialised."""
    h = None
    last = len(self.conv) - 1
    for i, (conv, bn) in enumerate(zip(self.conv, self.batch)):
        w = conv.weight.view(conv.out_channels, conv.in_channels)
        p = _dropout if i == last else 0.0                   # a following nn.Dropout, folded in (see UnitPointNet.forward)
        if i == 0:
            h = ops.linear_bn_act_cat_rows(rows1, rows2, w, conv.bias, bn, 0.0, p)
        else:
            h = ops.linear_bn_act_rows(h, w, conv.bias, bn, 0.0, p)
    return h


UnitPointNet.forward_rows_cat = _unit_forward_rows_cat


class SetAbstraction(nn.Module):
    """FPS -> ball-query group -> MiniPointNet -> max over K   [common.py:180-214]."""

    def __init__(self, C: int, radius: float, in_channels: int, mlps: list[int], K: int = 32,
                 pooling_type: str = 'max', grouping_norm: bool = False):
        super().__init__()
        self.point_net = MiniPointNet(in_channels, mlps)
        self.C = C
        self.radius = radius
        self.K = K
        self.pooling_type = pooling_type
        self.grouping_norm = grouping_norm
        self.fps_start = None          # optional (B,) first FPS pick (tests); None = reference's randint

    def forward(self, coords: torch.Tensor, features: torch.Tensor, _geom=None, lengths=None):
        """_geom (internal): (centroid coords, ball-query NeighborIndex) precomputed by ops.PyramidGeometry on the side
        stream; None computes them here, as the reference does.  lengths: real points per cloud (see sample())."""
        if _geom is not None:
            centroid_coords, nbr = _geom
            grouped = ops.group_points(coords, features, centroid_coords, nbr, self.radius if self.grouping_norm else None, pad4=True)
        else:
            centroid_coords = sample(coords, self.C, self.fps_start, lengths)
            grouped = _group_rows(centroid_coords, coords, features, self.radius, self.K, self.grouping_norm, lengths)
        if self.pooling_type == 'max':
            return centroid_coords, self.point_net.forward_rows(grouped, pool_max=True)   # (B,C,mlp[-1])
        x = self.point_net.forward_rows(grouped)             # point-major rows (B,C,K,*): no permute, no copy
        return centroid_coords, reduce(x, self.pooling_type)


class SetAbstractionMSG(nn.Module):
    """Multi-scale-grouping set abstraction (BASELINE configs[2]; the reference has no MSG class, SURVEY.md 8a-2: "MSG" =
    several `group` calls on ONE centroid set with different (r, K)).  Semantically
        centroids = sample(coords, C)
        out = cat([reduce(MiniPointNet_i(group(centroids, coords, features, r_i, K_i, grouping_norm)), pooling_type)], -1)
    with the reference's own pieces [common.py:6-91,125-150]; here FPS runs once, all ball queries share one scan of the
    points (ops.query_ball_point_multi) and every scale goes through the fused group -> MLP -> max path.
    Parameters: `point_nets.<i>.conv.<j>` / `point_nets.<i>.batch.<j>` (a ModuleList of the reference's MiniPointNet)."""

    def __init__(self, C: int, radii: list[float], in_channels: int, mlps_list: list[list[int]], Ks: list[int],
                 pooling_type: str = 'max', grouping_norm: bool = False):
        super().__init__()
        if not (len(radii) == len(mlps_list) == len(Ks)) or len(radii) == 0:
            raise ValueError("SetAbstractionMSG: radii, mlps_list and Ks must be non-empty lists of the same length")
        self.point_nets = nn.ModuleList(MiniPointNet(in_channels, mlps) for mlps in mlps_list)
        self.C = C
        self.radii = list(radii)
        self.Ks = list(Ks)
        self.pooling_type = pooling_type
        self.grouping_norm = grouping_norm
        self.fps_start = None          # optional (B,) first FPS pick (tests); None = reference's randint

    def forward(self, coords: torch.Tensor, features: torch.Tensor, _geom=None, lengths=None):
        """_geom (internal): (centroid coords, [NeighborIndex per scale]) precomputed by ops.PyramidGeometry on the side
        stream; None computes them here.  lengths: real points per cloud (see sample())."""
        if _geom is not None:
            centroid_coords, nbrs = _geom
        else:
            centroid_coords = sample(coords, self.C, self.fps_start, lengths)
            nbrs = [ops.NeighborIndex(idx, coords.shape[1])
                    for idx in ops.query_ball_point_multi(self.radii, self.Ks, coords, centroid_coords, lengths=lengths)]
        outs = []
        for r, nbr, net in zip(self.radii, nbrs, self.point_nets):
            grouped = ops.group_points(coords, features, centroid_coords, nbr, r if self.grouping_norm else None, pad4=True)
            if self.pooling_type == 'max':
                outs.append(net.forward_rows(grouped, pool_max=True))
            else:
                outs.append(reduce(net.forward_rows(grouped), self.pooling_type))
        return centroid_coords, torch.cat(outs, dim=-1)


class FeaturePropagation(nn.Module):
    """3-NN interpolation -> skip concat -> UnitPointNet   [common.py:217-243]."""

    def __init__(self, in_channels: int, mlps: list[int]):
        super().__init__()
        self.point_net = UnitPointNet(in_channels, mlps)

    def forward(self, coords_1, coords_2, features_1, features_2, _geom=None, lengths=None, _dropout: float = 0.0):
        """_geom (internal): (NeighborIndex, d2) of the 3-NN table precomputed by ops.PyramidGeometry.  lengths: real rows
        of coords_1 (see sample()).  _dropout (internal): see UnitPointNet.forward."""
        up = interpolate(features_2, coords_1, coords_2, lengths=lengths) if _geom is None else ops.three_interpolate(features_2, _geom[0], _geom[1])
        if features_1 is None:
            return self.point_net(up.permute(0, 2, 1), _dropout=_dropout).permute(0, 2, 1)
        return self.point_net.forward_rows_cat(features_1, up, _dropout)


class InvResMLP(nn.Module):
    """Inverted-residual block of PointNeXt: self ball query (always /r) -> MLP -> max -> MLP -> +x
    [common.py:246-301]."""

    def __init__(self, radius: int, in_channels: int, mlp_size: int, K: int, pooling_type: str = 'max'):
        super().__init__()
        self.radius = radius
        self.K = K
        self.pooling_type = pooling_type
        self.neighbour_features_mlp = MiniPointNet(in_channels, [mlp_size])
        self.point_features_mlp = UnitPointNet(mlp_size, [4 * mlp_size, mlp_size])

    def forward(self, centroid_coords, coords, features):
        grouped = _group_rows(centroid_coords, coords, features, self.radius, self.K, True)
        if self.pooling_type == 'max':
            x = self.neighbour_features_mlp.forward_rows(grouped, pool_max=True)
        else:
            x = reduce(self.neighbour_features_mlp.forward_rows(grouped), self.pooling_type)
        x = self.point_features_mlp(x.permute(0, 2, 1)).permute(0, 2, 1)
        return centroid_coords, x + features

"""Torch-CPU restatement of the reference hot path  -- TEST INFRASTRUCTURE ONLY.

This is the "what the reference computes" oracle: the same ATen op sequence the reference issues
(so the fp32 arithmetic is the reference's), restated independently, with two selection modes:

  tie="raw"    torch.topk, exactly as the reference (arbitrary order among equal keys)
  tie="canon"  stable sort -> lowest index wins among equal keys (the rule the CUDA path implements)

Also index-returning variants (the reference only returns gathered values) and an injectable FPS
start index (the reference draws it with torch.randint, models/utils/common.py:22).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Citations are relative to /root/reference/.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------- selection helper


def _select_smallest(keys: torch.Tensor, k: int, tie: str):
    """k smallest along the last dim, ascending.  Returns (values, indices int64)."""
    if tie == "raw":
        return torch.topk(keys, k, dim=-1, largest=False, sorted=True)
    order = torch.sort(keys, dim=-1, stable=True)
    return order.values[..., :k], order.indices[..., :k]


def _select_largest(keys: torch.Tensor, k: int, tie: str):
    if tie == "raw":
        return torch.topk(keys, k, dim=-1)
    order = torch.sort(keys, dim=-1, descending=True, stable=True)
    return order.values[..., :k], order.indices[..., :k]


def _rows(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """t (B,N,D), idx (B,...) int64 -> t[b, idx[b,...]]  (B,...,D)"""
    b = torch.arange(t.shape[0], device=idx.device).view(-1, *([1] * (idx.dim() - 1))).expand_as(idx)
    return t[b, idx]


# --------------------------------------------------------------------------- common.py ops


def fps_indices(coords: torch.Tensor, C: int, start: torch.Tensor | None = None) -> torch.Tensor:
    """Farthest point sampling, returns the picked INDICES (B,C) int32.

    Follows models/utils/common.py:17-31: running min of linalg.vector_norm distances, next pick =
    torch.max (lowest index on ties).  `start` replaces the randint draw of common.py:22; when None
    the same draw is made (same shape/dtype -> same generator consumption)."""
    B, N, _ = coords.shape
    if start is None:
        start = torch.randint(0, N, (B,), dtype=torch.int, device=coords.device)
    coords = coords.float()          # selections are always made in the reference's fp32 arithmetic
    far = start.to(torch.int64)
    rows = torch.arange(B, device=coords.device)
    picks = torch.zeros(B, C, dtype=torch.int32)                 # a CPU tensor, as common.py:20 (one D2H copy per pick on CUDA)
    running = torch.full((B, N), torch.inf, device=coords.device)
    for i in range(C):
        picks[:, i] = far.to(torch.int32)
        centre = coords[rows, far, :].view(B, 1, 3)
        d = torch.linalg.vector_norm(coords - centre, dim=-1)
        running = torch.where(d < running, d, running)
        far = torch.max(running, -1)[1]
    return picks


def sample(coords: torch.Tensor, C: int, start: torch.Tensor | None = None) -> torch.Tensor:
    """models/utils/common.py:6-34 -> coordinates (B,C,3) of the FPS picks."""
    return _rows(coords, fps_indices(coords, C, start).long().to(coords.device))


def ball_query_indices(centroids, coords, r: float, K: int, tie: str = "canon") -> torch.Tensor:
    """common.py:54-61 -> (B,C,K) int64.  (fp32 arithmetic even when the features run in fp64.)"""
    diff = coords.float().unsqueeze(1) - centroids.float().unsqueeze(2)          # (B,C,N,3) points - centroids
    d2 = (diff ** 2).sum(dim=-1)
    d2 = torch.where(d2 <= r ** 2, d2, torch.full_like(d2, torch.inf))
    return _select_smallest(d2, K, tie)[1]


def group(centroids, coords, features, r: float, K: int, normalize: bool = False, tie: str = "canon",
          idx: torch.Tensor | None = None) -> torch.Tensor:
    """common.py:37-71 -> (B,C,K,3+D)."""
    if idx is None:
        idx = ball_query_indices(centroids, coords, r, K, tie)
    local = _rows(coords, idx) - centroids.unsqueeze(2)
    if normalize:
        local = local / r
    return torch.cat([local, _rows(features, idx)], dim=-1)


def reduce(x: torch.Tensor, type: str) -> torch.Tensor:
    """common.py:74-91 (the 'avg' branch keeps the reference's `[0]` quirk, common.py:89)."""
    if type == "max":
        return torch.max(x, dim=2)[0]
    if type == "avg":
        return torch.mean(x, dim=2)[0]
    raise ValueError(f"'{type}' pooling not supported; use 'max' or 'avg'.")


def three_nn(coords_1, coords_2, k: int = 3, tie: str = "canon"):
    """common.py:110-114 -> (d2 (B,N,k), idx (B,N,k) int64); coords_1 are the queries."""
    diff = coords_2.float().unsqueeze(1) - coords_1.float().unsqueeze(2)         # (B,N,M,3)
    d2 = (diff ** 2).sum(dim=-1)
    vals, idx = _select_smallest(d2, k, tie)
    return vals.to(coords_1.dtype), idx


def interpolate(points, coords_1, coords_2, k: int = 3, tie: str = "canon") -> torch.Tensor:
    """common.py:94-122 -> (B,N,D)."""
    d2, idx = three_nn(coords_1, coords_2, k, tie)
    nbr = _rows(points, idx)                                     # (B,N,k,D)
    w = 1.0 / (d2.unsqueeze(-1) + 1e-9)
    norm = torch.sum(w, dim=2, keepdim=True)
    return torch.sum(nbr * w / norm, dim=2)


# --------------------------------------------------------------------------- dgcnn.py ops


def pairwise_neg_sqdist(x: torch.Tensor) -> torch.Tensor:
    """dgcnn.py:16-18: x (B,F,N) -> (B,N,N) negative squared distances in the reference's form."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return -xx - inner - xx.transpose(2, 1)


def knn(x: torch.Tensor, k: int, tie: str = "canon") -> torch.Tensor:
    """dgcnn.py:7-21 -> (B,N,k) int64."""
    return _select_largest(pairwise_neg_sqdist(x), k, tie)[1]


def get_graph_feature(x: torch.Tensor, k: int = 20, idx: torch.Tensor | None = None, tie: str = "canon"):
    """dgcnn.py:24-57 (dim9=False) -> (B,2F,N,k).  Device taken from x (the reference's
    torch.cuda.is_available() device pick, dgcnn.py:39, is a bug we do not restate)."""
    B, Fd, N = x.shape
    if idx is None:
        idx = knn(x, k, tie)
    pts = x.transpose(2, 1).contiguous()                         # (B,N,F)
    nbr = _rows(pts, idx)                                        # (B,N,k,F)
    ctr = pts.unsqueeze(2).expand(B, N, idx.shape[-1], Fd)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2).contiguous()


# --------------------------------------------------------------------------- modules (same parameter names)


class MiniPointNet(nn.Module):
    """common.py:125-150"""

    def __init__(self, in_channels, mlps):
        super().__init__()
        self.conv = nn.ModuleList()
        self.batch = nn.ModuleList()
        c = in_channels
        for w in mlps:
            self.conv.append(nn.Conv2d(c, w, (1, 1)))
            self.batch.append(nn.BatchNorm2d(w))
            c = w

    def forward(self, x):
        for conv, bn in zip(self.conv, self.batch):
            x = F.relu(bn(conv(x)))
        return x


class UnitPointNet(nn.Module):
    """common.py:153-178"""

    def __init__(self, in_channels, mlps):
        super().__init__()
        self.conv = nn.ModuleList()
        self.batch = nn.ModuleList()
        c = in_channels
        for w in mlps:
            self.conv.append(nn.Conv1d(c, w, 1))
            self.batch.append(nn.BatchNorm1d(w))
            c = w

    def forward(self, x):
        for conv, bn in zip(self.conv, self.batch):
            x = F.relu(bn(conv(x)))
        return x


class SetAbstraction(nn.Module):
    """common.py:180-214; `tie` and `fps_start` are oracle-only knobs."""

    def __init__(self, C, radius, in_channels, mlps, K=32, pooling_type="max", grouping_norm=False, tie="canon"):
        super().__init__()
        self.point_net = MiniPointNet(in_channels, mlps)
        self.C, self.radius, self.K = C, radius, K
        self.pooling_type, self.grouping_norm, self.tie = pooling_type, grouping_norm, tie
        self.fps_start = None

    def forward(self, coords, features):
        cen = sample(coords, self.C, self.fps_start)
        g = group(cen, coords, features, self.radius, self.K, self.grouping_norm, self.tie)
        g = self.point_net(g.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        return cen, reduce(g, self.pooling_type)


class FeaturePropagation(nn.Module):
    """common.py:217-243"""

    def __init__(self, in_channels, mlps, tie="canon"):
        super().__init__()
        self.point_net = UnitPointNet(in_channels, mlps)
        self.tie = tie

    def forward(self, coords_1, coords_2, features_1, features_2):
        up = interpolate(features_2, coords_1, coords_2, tie=self.tie)
        feats = up if features_1 is None else torch.cat([features_1, up], dim=-1)
        return self.point_net(feats.permute(0, 2, 1)).permute(0, 2, 1)


class InvResMLP(nn.Module):
    """common.py:246-301"""

    def __init__(self, radius, in_channels, mlp_size, K, pooling_type="max", tie="canon"):
        super().__init__()
        self.radius, self.K, self.pooling_type, self.tie = radius, K, pooling_type, tie
        self.neighbour_features_mlp = MiniPointNet(in_channels, [mlp_size])
        self.point_features_mlp = UnitPointNet(mlp_size, [4 * mlp_size, mlp_size])

    def forward(self, centroid_coords, coords, features):
        g = group(centroid_coords, coords, features, self.radius, self.K, True, self.tie)
        g = self.neighbour_features_mlp(g.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        f = reduce(g, self.pooling_type).permute(0, 2, 1)
        f = self.point_features_mlp(f).permute(0, 2, 1)
        return centroid_coords, f + features


class PointNetpp(nn.Module):
    """models/PointNetpp/PointNetpp.py:6-48 (SSG semseg)."""

    def __init__(self, part_classes, tie="canon"):
        super().__init__()
        self.sa1 = SetAbstraction(1024, 0.1, 6 + 3, [32, 32, 64], tie=tie)
        self.sa2 = SetAbstraction(256, 0.2, 64 + 3, [64, 64, 128], tie=tie)
        self.sa3 = SetAbstraction(64, 0.4, 128 + 3, [128, 128, 256], tie=tie)
        self.sa4 = SetAbstraction(16, 0.8, 256 + 3, [256, 256, 512], tie=tie)
        self.fp4 = FeaturePropagation(768, [256, 256], tie=tie)
        self.fp3 = FeaturePropagation(384, [256, 256], tie=tie)
        self.fp2 = FeaturePropagation(320, [256, 128], tie=tie)
        self.fp1 = FeaturePropagation(128, [128, 128, 128, 128], tie=tie)
        self.drop = nn.Dropout(0.5)
        self.conv = nn.Conv1d(128, part_classes, 1)

    def forward(self, x):
        c0, f0 = x[:, :, :3], x[:, :, 3:]
        c1, f1 = self.sa1(c0, f0)
        c2, f2 = self.sa2(c1, f1)
        c3, f3 = self.sa3(c2, f2)
        c4, f4 = self.sa4(c3, f3)
        f3 = self.fp4(c3, c4, f3, f4)
        f2 = self.fp3(c2, c3, f2, f3)
        f1 = self.fp2(c1, c2, f1, f2)
        f0 = self.fp1(c0, c1, None, f1)
        y = self.conv(self.drop(f0).permute(0, 2, 1))
        return y.permute(0, 2, 1)


class SetAbstractionMSG(nn.Module):
    """Multi-scale grouping composed from the reference's pieces ("MSG" = several `group` calls on one
    centroid set): sample once (common.py:6-34), then per scale group -> MiniPointNet -> reduce (common.py:37-91,
    125-150, 205-214), concatenated along the channels."""

    def __init__(self, C, radii, in_channels, mlps_list, Ks, pooling_type="max", grouping_norm=False, tie="canon"):
        super().__init__()
        self.point_nets = nn.ModuleList(MiniPointNet(in_channels, m) for m in mlps_list)
        self.C, self.radii, self.Ks = C, list(radii), list(Ks)
        self.pooling_type, self.grouping_norm, self.tie = pooling_type, grouping_norm, tie
        self.fps_start = None

    def forward(self, coords, features):
        cen = sample(coords, self.C, self.fps_start)
        outs = []
        for r, K, net in zip(self.radii, self.Ks, self.point_nets):
            g = group(cen, coords, features, r, K, self.grouping_norm, self.tie)
            g = net(g.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
            outs.append(reduce(g, self.pooling_type))
        return cen, torch.cat(outs, dim=-1)


class PointNetppMSG(nn.Module):
    """The SSG skeleton of models/PointNetpp/PointNetpp.py:6-48 with multi-scale set abstractions (BASELINE configs[2])."""

    def __init__(self, part_classes, tie="canon"):
        super().__init__()
        self.sa1 = SetAbstractionMSG(1024, [0.05, 0.1], 9, [[16, 16, 32], [32, 32, 64]], [16, 32], tie=tie)
        self.sa2 = SetAbstractionMSG(256, [0.1, 0.2], 96 + 3, [[64, 64, 128], [64, 64, 128]], [16, 32], tie=tie)
        self.sa3 = SetAbstractionMSG(64, [0.2, 0.4], 256 + 3, [[128, 128, 256], [128, 128, 256]], [16, 32], tie=tie)
        self.sa4 = SetAbstractionMSG(16, [0.4, 0.8], 512 + 3, [[256, 256, 512], [256, 256, 512]], [16, 32], tie=tie)
        self.fp4 = FeaturePropagation(1024 + 512, [256, 256], tie=tie)
        self.fp3 = FeaturePropagation(256 + 256, [256, 256], tie=tie)
        self.fp2 = FeaturePropagation(256 + 96, [256, 128], tie=tie)
        self.fp1 = FeaturePropagation(128, [128, 128, 128, 128], tie=tie)
        self.drop = nn.Dropout(0.5)
        self.conv = nn.Conv1d(128, part_classes, 1)

    forward = PointNetpp.forward


class PointNeXt(nn.Module):
    """models/PointNeXt/PointNeXt.py:17-147 (`version` is unused there too, :22)."""

    def __init__(self, part_classes, version="b", tie="canon"):
        super().__init__()
        self.num_classes = part_classes
        self.mlp = UnitPointNet(9, [32])
        self.sa1 = SetAbstraction(1024, 0.1, 32 + 3, [32, 32, 64], grouping_norm=True, tie=tie)
        self.irmlp1 = InvResMLP(0.1, 64 + 3, 64, 32, tie=tie)
        self.sa2 = SetAbstraction(256, 0.2, 64 + 3, [64, 64, 128], grouping_norm=True, tie=tie)
        self.irmlp2 = InvResMLP(0.1, 128 + 3, 128, 32, tie=tie)
        self.irmlp2_1 = InvResMLP(0.2, 128 + 3, 128, 32, tie=tie)
        self.sa3 = SetAbstraction(64, 0.4, 128 + 3, [128, 128, 256], grouping_norm=True, tie=tie)
        self.irmlp3 = InvResMLP(0.4, 256 + 3, 256, 32, tie=tie)
        self.sa4 = SetAbstraction(16, 0.8, 256 + 3, [256, 256, 512], grouping_norm=True, tie=tie)
        self.irmlp4 = InvResMLP(0.8, 512 + 3, 512, 16, tie=tie)
        self.fp4 = FeaturePropagation(512 + 256, [256, 256], tie=tie)
        self.fp3 = FeaturePropagation(256 + 128, [256, 256], tie=tie)
        self.fp2 = FeaturePropagation(256 + 64, [256, 128], tie=tie)
        self.fp1 = FeaturePropagation(128 + 32, [128, 128, 128, 128], tie=tie)
        self.drop = nn.Dropout(0.5)
        self.conv = nn.Conv1d(128, part_classes, 1)

    def forward(self, x):
        xt = x.permute(0, 2, 1)
        c0 = xt[:, :3, :].permute(0, 2, 1)
        f0 = self.mlp(xt).permute(0, 2, 1)
        c1, f1 = self.sa1(c0, f0)
        c1, f1 = self.irmlp1(c1, c1, f1)
        c2, f2 = self.sa2(c1, f1)
        c2, f2 = self.irmlp2(c2, c2, f2)
        c2, f2 = self.irmlp2_1(c2, c2, f2)
        c3, f3 = self.sa3(c2, f2)
        c3, f3 = self.irmlp3(c3, c3, f3)
        c4, f4 = self.sa4(c3, f3)
        c4, f4 = self.irmlp4(c4, c4, f4)
        f3 = self.fp4(c3, c4, f3, f4)
        f2 = self.fp3(c2, c3, f2, f3)
        f1 = self.fp2(c1, c2, f1, f2)
        f0 = self.fp1(c0, c1, f0, f1)
        y = self.conv(self.drop(f0).permute(0, 2, 1))
        return y.permute(0, 2, 1)


class EdgeConv(nn.Module):
    """dgcnn.py:60-77"""

    def __init__(self, in_channels, out_channels, k=20, tie="canon"):
        super().__init__()
        self.k, self.tie = k, tie
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels * 2, out_channels, kernel_size=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.LeakyReLU(negative_slope=0.2),
        )

    def forward(self, x):
        return self.conv(get_graph_feature(x, self.k, tie=self.tie)).max(dim=-1)[0]


def _head(cin, cout, drop=None):
    layers = [nn.Conv1d(cin, cout, kernel_size=1, bias=False), nn.BatchNorm1d(cout), nn.LeakyReLU(negative_slope=0.2)]
    if drop is not None:
        layers.append(nn.Dropout(drop))
    return nn.Sequential(*layers)


class DGCNN(nn.Module):
    """dgcnn.py:80-162"""

    def __init__(self, num_classes=13, k=20, emb_dims=1024, dropout=0.5, tie="canon"):
        super().__init__()
        self.k, self.num_classes = k, num_classes
        self.conv1 = EdgeConv(3, 64, k, tie)
        self.conv2 = EdgeConv(64, 64, k, tie)
        self.conv3 = EdgeConv(64, 64, k, tie)
        self.conv4 = EdgeConv(64, 128, k, tie)
        self.conv5 = _head(320, emb_dims)
        self.conv6 = _head(emb_dims + 320, 512, dropout)
        self.conv7 = _head(512, 256, dropout)
        self.conv8 = nn.Conv1d(256, num_classes, kernel_size=1)

    def forward(self, x):
        xyz = x[:, :3, :] if x.size(1) == 6 else x
        x1 = self.conv1(xyz)
        x2 = self.conv2(x1)
        x3 = self.conv3(x2)
        x4 = self.conv4(x3)
        cat = torch.cat((x1, x2, x3, x4), dim=1)
        x5 = self.conv5(cat)
        y = self.conv8(self.conv7(self.conv6(torch.cat((cat, x5), dim=1))))
        return y.transpose(2, 1).contiguous(), x5, None


class DGCNNWithColor(nn.Module):
    """dgcnn.py:165-257"""

    def __init__(self, num_classes=13, k=20, emb_dims=1024, dropout=0.5, tie="canon"):
        super().__init__()
        self.k, self.num_classes = k, num_classes
        self.conv1 = EdgeConv(3, 64, k, tie)
        self.conv2 = EdgeConv(64, 64, k, tie)
        self.conv3 = EdgeConv(64, 64, k, tie)
        self.conv4 = EdgeConv(64, 128, k, tie)
        self.color_conv = _head(3, 64)
        self.conv5 = _head(384, emb_dims)
        self.conv6 = _head(emb_dims + 384, 512, dropout)
        self.conv7 = _head(512, 256, dropout)
        self.conv8 = nn.Conv1d(256, num_classes, kernel_size=1)

    def forward(self, x):
        if x.size(1) != 6:
            raise ValueError("DGCNNWithColor expects 6-channel input (xyz + rgb)")
        x1 = self.conv1(x[:, :3, :])
        x2 = self.conv2(x1)
        x3 = self.conv3(x2)
        x4 = self.conv4(x3)
        cat = torch.cat((x1, x2, x3, x4, self.color_conv(x[:, 3:6, :])), dim=1)
        x5 = self.conv5(cat)
        y = self.conv8(self.conv7(self.conv6(torch.cat((cat, x5), dim=1))))
        return y.transpose(2, 1).contiguous(), x5, None


# --------------------------------------------------------------------------- synthetic S3DIS-shaped data


def s3dis_blocks(B: int, N: int = 4096, seed: int = 0, classes: int = 13):
    """SURVEY.md §8(d) generator: (points (B,N,9) f32, labels (B,N,classes) u8, lengths (B,) i64).
    xyz = block origin + U[0,1)^2 x 3U[0,1); rgb 0..255; ch6:9 = xyz - block centre
    (data_processing/preprocess_dataset.py:73-90)."""
    g = torch.Generator().manual_seed(seed)
    origin = torch.randint(0, 20, (B, 1, 2), generator=g).float()
    xy = origin + torch.rand(B, N, 2, generator=g)
    z = 3.0 * torch.rand(B, N, 1, generator=g)
    xyz = torch.cat([xy, z], dim=-1)
    rgb = torch.randint(0, 256, (B, N, 3), generator=g).float()
    zc = (z.amin(dim=1, keepdim=True) + z.amax(dim=1, keepdim=True)) / 2
    centre = torch.cat([origin + 0.5, zc], dim=-1)
    pts = torch.cat([xyz, rgb, xyz - centre], dim=-1)
    lab = F.one_hot(torch.randint(0, classes, (B, N), generator=g), classes).to(torch.uint8)
    return pts, lab, torch.full((B,), N, dtype=torch.int64)


# --------------------------------------------------------------------------- Training/metrics.py


def metrics_confusion_matrix(predictions: torch.Tensor, labels: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Restatement of Training/metrics.py:52-78: matrix[i, j] = #points with label i predicted j, unpadded points only
    (torch.argmax: first maximum wins)."""
    B, _, C = labels.shape
    m = torch.zeros(C, C, dtype=torch.int64)
    for b in range(B):
        n = int(mask[b])
        pc = predictions[b, :n].argmax(-1)
        lc = labels[b, :n].argmax(-1)
        m += torch.bincount(lc * C + pc, minlength=C * C).view(C, C)
    return m


def metrics_update_accuracy(predictions, labels, mask):
    """Training/metrics.py:28-50 -> (correct, total)."""
    m = metrics_confusion_matrix(predictions, labels, mask)
    return int(m.diagonal().sum()), int(mask.sum())


def metrics_update_iou(predictions, labels, mask):
    """Training/metrics.py:115-146 -> (intersections (C,), unions (C,)) float32.  The reference tests
    `labels[..., c] == 1` (:137), NOT the label argmax the confusion matrix uses: a row without any 1 (malformed, or a
    padding row inside `mask`) belongs to no class here, and only enlarges the union of the class it is predicted as."""
    B, _, C = labels.shape
    inter, union = torch.zeros(C), torch.zeros(C)
    for b in range(B):
        n = int(mask[b])
        pc = predictions[b, :n].argmax(-1)
        for c in range(C):
            lm = labels[b, :n, c] == 1
            pm = pc == c
            inter[c] += float((lm & pm).sum())
            union[c] += float((lm | pm).sum())
    return inter, union


def metrics_iou(predictions, labels, mask):
    """Training/metrics.py:81-112 -> (mean IoU, per-class IoU) with eps = 1e-6."""
    inter, union = metrics_update_iou(predictions, labels, mask)
    ious = ((inter.double() + 1e-6) / (union.double() + 1e-6)).to(torch.float32)
    return ious.mean().item(), ious


# --------------------------------------------------------------------------- block dataloader


def block_getitem(points: torch.Tensor, labels: torch.Tensor, sampling: int | None):
    """data_processing/block_datasets.py:117-128 for one loaded block: the random row selection (host generator)."""
    if sampling is not None:
        n = points.shape[0]
        rows = torch.randperm(n)[:sampling] if n > sampling else torch.randint(n, (sampling,))
        points, labels = points[rows], labels[rows]
    return points, labels


def collate_blocks(batch):
    """block_datasets.py:5-29 -> (points (B,N,9) f32, labels (B,N,L) u8, lengths (B,) int64), zero padded."""
    B, N = len(batch), max(p.shape[0] for p, _ in batch)
    pts = torch.zeros(B, N, batch[0][0].shape[1], dtype=torch.float32)
    lab = torch.zeros(B, N, batch[0][1].shape[1], dtype=torch.uint8)
    for b, (p, l) in enumerate(batch):
        pts[b, :p.shape[0]] = p
        lab[b, :p.shape[0]] = l
    return pts, lab, torch.tensor([p.shape[0] for p, _ in batch], dtype=torch.int64)


def gather_block_batch(blocks, block_ids, sel):
    """The batch a given plan produces: rows sel[b] of block block_ids[b] (sel None: all rows, zero padded)."""
    if sel is None:
        return collate_blocks([blocks[i] for i in block_ids])
    return collate_blocks([(blocks[i][0][sel[b].long()], blocks[i][1][sel[b].long()]) for b, i in enumerate(block_ids)])


# --------------------------------------------------------------------------- sliding-window inference


def predict_single_scene(model, points: torch.Tensor, batch_size: int = 4096, overlap: int = 512):
    """models/dgcnn/utils.py:67-131: window after window through `model` ((1,F,n) -> (logits (1,n,C), _, _)),
    overlap-add, divide by the cover count, argmax and max softmax.  -> (mean_logits (N,C), predictions (N,), conf (N,));
    the reference returns the last two."""
    n = points.shape[0]
    with torch.no_grad():
        if n <= batch_size:
            mean = model(points.T.unsqueeze(0))[0].squeeze(0)
        else:
            step = batch_size - overlap
            total = torch.zeros(n, model.num_classes, device=points.device)
            count = torch.zeros(n, device=points.device)
            for start in range(0, n, step):
                end = min(start + batch_size, n)
                total[start:end] += model(points[start:end].T.unsqueeze(0))[0].squeeze(0)
                count[start:end] += 1
            mean = total / count.unsqueeze(1)
    return mean, torch.argmax(mean, dim=1), torch.softmax(mean, dim=1).max(dim=1)[0]

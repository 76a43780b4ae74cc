"""Host-side mirror of the reference's models/dgcnn/dgcnn.py on top of libpcnbr.

knn / get_graph_feature / EdgeConv / DGCNN / DGCNNWithColor / get_model / get_loss keep the
reference's signatures, return layouts and state_dict keys (/root/reference/models/dgcnn/dgcnn.py).
Differences, all deliberate: the device comes from the input tensor (the reference asks
torch.cuda.is_available(), dgcnn.py:39); ties in the k-NN selection go to the lowest index; the
(B,N,N) distance matrix and the three (B,N,k,F)-sized temporaries are never materialised.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import ops

__all__ = ["knn", "get_graph_feature", "EdgeConv", "DGCNN", "DGCNNWithColor", "get_model", "get_loss"]


def knn(x: torch.Tensor, k: int, lengths=None) -> torch.Tensor:
    """x (B,F,N) -> LongTensor (B,N,k) of the k nearest points in feature space, self first
    [dgcnn.py:7-21].  lengths (B,): length-aware form for zero-padded batches -- rows n < lengths[b] get
    the reference's result on the cloud passed alone, the padding rows a filler."""
    return ops.knn_graph(x, k, lengths=lengths).long()


def _point_major(x: torch.Tensor) -> torch.Tensor:
    """(B,F,N) -> (B,N,F) contiguous; free when x is already a transposed view of point-major memory."""
    return x.transpose(2, 1).contiguous()


def get_graph_feature(x: torch.Tensor, k: int = 20, idx: torch.Tensor | None = None, dim9: bool = False):
    """x (B,F,N) -> edge features (B,2F,N,k) = cat(x_j - x_i, x_i) [(B,3F,N,k) when dim9]
    [dgcnn.py:24-57].  The result is laid out point-major in memory (channels-last)."""
    B, N = x.size(0), x.size(2)
    x = x.reshape(B, -1, N)
    if idx is None:
        idx32 = ops.knn_graph(x if not dim9 else x[:, 6:], k)
    else:
        idx32 = ops._as_i32(idx.reshape(B, N, -1))
    nbr = ops.NeighborIndex(idx32, N)
    out = ops.edge_features(_point_major(x), nbr).permute(0, 3, 1, 2)
    if dim9:
        out = torch.cat((out, out[:, x.size(1):]), dim=1)
    return out


class EdgeConv(nn.Module):
    """get_graph_feature -> Conv2d 1x1 (no bias) -> BatchNorm2d -> LeakyReLU(0.2) -> max over k
    [dgcnn.py:60-77]."""

    def __init__(self, in_channels, out_channels, k=20):
        super().__init__()
        self.k = k
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels * 2, out_channels, kernel_size=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.LeakyReLU(negative_slope=0.2),
        )
        # fused = algebraic split W.[xj-xi; xi] = A.xj + (B-A).xi: no (B,*,N,k) tensor is ever built
        # (SURVEY.md 8f-2).  fused=False runs the reference's op sequence literally on the K9/K6 kernels.
        self.fused = os.environ.get("PCNBR_EDGECONV_EXACT") is None and out_channels in (32, 64, 128, 256)

    def forward(self, x, _nbr=None, lengths=None):
        """_nbr (internal): the layer's kNN table (ops.NeighborIndex) when it was computed ahead of time -- only the first
        layer's graph depends on the input coordinates alone (DGCNN.prepare_geometry).  lengths: see knn()."""
        if not self.fused:
            if _nbr is None and lengths is not None:
                _nbr = ops.NeighborIndex(ops.knn_graph(x, self.k, lengths=lengths), x.shape[2])
            x = get_graph_feature(x, k=self.k, idx=_nbr.idx if _nbr is not None else None)
            x = self.conv(x)
            return ops.max_pool_neighbors(x, -1)
        conv, bn, act = self.conv[0], self.conv[1], self.conv[2]
        B, F, N = x.shape
        O = conv.out_channels
        nbr = _nbr if _nbr is not None else ops.NeighborIndex(ops.knn_graph(x, self.k, lengths=lengths), N)
        W = conv.weight.view(O, 2 * F)
        Wcat = _SplitWeightFn.apply(W)                                      # [A ; B - A]  (2O, F)
        rows = _point_major(x)
        if F % 4:
            # xyz layer.  u_ij = A (x_j - x_i) + B x_i is evaluated as P_j + Q_i with P = A x, Q = (B - A) x, which cancels
            # catastrophically when the cloud sits tens of metres from the origin (S3DIS coordinates).  A is translation
            # invariant, so work on coordinates shifted per cloud, xc = x - m:  P = A xc,  Q = (B - A) xc + B m.  The rows
            # are then zero-padded 3 -> 4 channels so that all three GEMMs (incl. the 65536-row weight gradient) run on
            # tcgen05.
            # Length-aware evaluation: the shift is the cloud's first point instead (any point of the cloud does; unlike the
            # mean it does not see the zero padding and is the same whether the cloud comes alone or inside a padded batch).
            m = rows.mean(dim=1, keepdim=True) if lengths is None else rows[:, :1, :]       # (B,1,F)
            PQ = ops.linear_rows(_pad4(rows - m), _pad4(Wcat), None)
            PQ = PQ + torch.nn.functional.pad(torch.matmul(m, W[:, F:].t()), (O, 0))        # [0 | B m] broadcast over the points
        else:
            PQ = ops.linear_rows(rows, Wcat, None)                          # (B,N,F) x (F,2O): tensor-core GEMM / split-K wgrad
        out = ops.edgeconv_fused(PQ, nbr, bn, act.negative_slope)           # (B,N,O), point-major
        return out.permute(0, 2, 1)                                         # (B,O,N) view, no copy


class _SplitWeightFn(torch.autograd.Function):
    """W = [A | B] (O, 2F) -> [A ; B - A] (2O, F), the weights of the algebraic split, as ONE autograd node: its backward is
    gW = [gP - gQ | gQ] in two small kernels (slicing + subtraction + concatenation through autograd were ten launches of
    1-2 us per EdgeConv layer and pass)."""

    @staticmethod
    def forward(ctx, W):
        F = W.shape[1] // 2
        return torch.cat((W[:, :F], W[:, F:] - W[:, :F]), dim=0)

    @staticmethod
    def backward(ctx, g):
        O = g.shape[0] // 2
        return torch.cat((g[:O] - g[O:], g[O:]), dim=1)


def _pad4(t: torch.Tensor) -> torch.Tensor:
    """Zero-pad the last (channel) dimension to a multiple of 4 floats: the 16-byte row pitch TMA needs.  Padding both
    the rows and the weight leaves the product and every gradient unchanged (the pad columns are sliced away by
    autograd)."""
    return torch.nn.functional.pad(t, (0, (-t.shape[-1]) % 4))


def _fusable_dropout(mods):
    """(p, remaining modules): a leading nn.Dropout is folded into the fused conv+BatchNorm+activation node (p in training
    mode; in eval mode or with p = 0 it is the identity and simply dropped)."""
    if mods and isinstance(mods[0], nn.Dropout):
        d = mods[0]
        if not d.training or d.p == 0.0:
            return 0.0, mods[1:]
        if d.p < 1.0:
            return float(d.p), mods[1:]
    return 0.0, mods


def _run_pointwise(seq, rows: torch.Tensor) -> torch.Tensor:
    """seq = [Conv1d(kernel 1), BatchNorm1d, LeakyReLU(, Dropout)] applied to point-major rows (B,N,Cin) -> (B,N,Cout).
    A 1x1 convolution IS a GEMM: it is issued as ONE GEMM over the B*N rows (no transposes, no per-batch GEMMs) instead
    of cuDNN's fp32 convolution engines; BatchNorm1d + LeakyReLU run as the fused row kernels over the same rows
    (identical statistics).  Same parameters, same math."""
    from .common import _batch_norm_rows
    mods = list(seq) if isinstance(seq, nn.Sequential) else [seq]
    conv = mods[0]
    w = conv.weight.squeeze(-1)
    if w.shape[1] % 4 and w.shape[1] <= 16:                                 # the 3-channel colour branch
        rows, w = _pad4(rows), _pad4(w)
    if len(mods) >= 3 and isinstance(mods[1], nn.modules.batchnorm._BatchNorm) and isinstance(mods[2], nn.LeakyReLU):
        p_drop, rest = _fusable_dropout(mods[3:])
        y = ops.linear_bn_act_rows(rows, w, conv.bias, mods[1], mods[2].negative_slope, p_drop)
    else:
        y = ops.linear_rows(rows, w, conv.bias)
        rest = mods[1:]
    for m in rest:
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            y = _batch_norm_rows(m, y.view(-1, y.shape[-1])).view(y.shape)
        else:
            y = m(y)
    return y


def _run_pointwise_cat(seq, rows1: torch.Tensor, rows2: torch.Tensor) -> torch.Tensor:
    """_run_pointwise(seq, cat((rows1, rows2), dim=2)) without materialising the concatenation (the (B,N,1408) input of
    conv6 is 369 MB at 16 x 4096 points): the GEMM reads its K blocks from the two tensors in turn."""
    mods = list(seq) if isinstance(seq, nn.Sequential) else [seq]
    conv = mods[0]
    if not (len(mods) >= 3 and isinstance(mods[1], nn.modules.batchnorm._BatchNorm) and isinstance(mods[2], nn.LeakyReLU)):
        return _run_pointwise(seq, torch.cat((rows1, rows2), dim=2))
    p_drop, rest = _fusable_dropout(mods[3:])
    y = ops.linear_bn_act_cat_rows(rows1, rows2, conv.weight.squeeze(-1), conv.bias, mods[1], mods[2].negative_slope, p_drop)
    for m in rest:
        y = m(y)
    return y


def _first_layer_graph(xyz: torch.Tensor, k: int, stream=None):
    """The kNN table of the first EdgeConv (dgcnn.py:73-74 on the input coordinates) with its CSR inverse, as a list of
    tensors [idx, offsets, perm] (+ a trailing keep-alive object): the only graph of the network that depends on the input
    alone, so a training loop can compute it for the NEXT batch on the side stream (train.GraphedTrainStep).  stream: run
    the kernels there, after the torch-side preparation on the current stream (the caller waits on it before use)."""
    xc = xyz.contiguous()
    if stream is not None:
        stream.wait_event(torch.cuda.current_stream().record_event())
    keep = [xc]
    with (ops.on_stream(stream) if stream is not None else ops._NullCtx()):
        nbr = ops.NeighborIndex(ops.knn_graph(xc, k, _keep=keep), xc.shape[2])
        off, perm, ws = nbr._build(ops._stream())
    nbr._csr = (off, perm)
    return [nbr.idx, off, perm, (nbr, ws, keep)]


def _graph_from(geometry, n_src: int):
    if geometry is None:
        return None
    nbr = ops.NeighborIndex(geometry[0], n_src)
    nbr._csr = (geometry[1], geometry[2])
    return nbr


def _pointwise(cin, cout, dropout=None):
    layers = [nn.Conv1d(cin, cout, kernel_size=1, bias=False), nn.BatchNorm1d(cout), nn.LeakyReLU(negative_slope=0.2)]
    if dropout is not None:
        layers.append(nn.Dropout(dropout))
    return nn.Sequential(*layers)


class DGCNN(nn.Module):
    """DGCNN semantic segmentation on xyz   [dgcnn.py:80-162].  Returns (logits (B,N,cls), x5, None)."""

    def __init__(self, num_classes=13, k=20, emb_dims=1024, dropout=0.5):
        super().__init__()
        self.k = k
        self.num_classes = num_classes
        self.conv1 = EdgeConv(3, 64, k)
        self.conv2 = EdgeConv(64, 64, k)
        self.conv3 = EdgeConv(64, 64, k)
        self.conv4 = EdgeConv(64, 128, k)
        self.conv5 = _pointwise(320, emb_dims)
        self.conv6 = _pointwise(emb_dims + 320, 512, dropout)
        self.conv7 = _pointwise(512, 256, dropout)
        self.conv8 = nn.Conv1d(256, num_classes, kernel_size=1)

    def prepare_geometry(self, x, stream=None):
        return _first_layer_graph(x[:, :3, :] if x.size(1) == 6 else x, self.k, stream)

    def forward(self, x, geometry=None, lengths=None):
        """lengths (B,): length-aware evaluation of a zero-padded batch: every kNN graph is built among the
        real points of each cloud; in eval mode the logits of the real rows are the reference's on the cloud passed alone."""
        xyz = x[:, :3, :] if x.size(1) == 6 else x
        x1 = self.conv1(xyz, _nbr=_graph_from(geometry, xyz.shape[2]), lengths=lengths)
        x2 = self.conv2(x1, lengths=lengths)
        x3 = self.conv3(x2, lengths=lengths)
        x4 = self.conv4(x3, lengths=lengths)
        # the head runs point-major: (B,N,C) rows, one GEMM per layer, logits come out as (B,N,classes)
        r_cat = torch.cat([t.permute(0, 2, 1) for t in (x1, x2, x3, x4)], dim=2)
        r5 = _run_pointwise(self.conv5, r_cat)
        r7 = _run_pointwise(self.conv7, _run_pointwise_cat(self.conv6, r_cat, r5))
        logits = _run_pointwise(self.conv8, r7)
        return logits, r5.permute(0, 2, 1), None


class DGCNNWithColor(nn.Module):
    """DGCNN with an RGB branch   [dgcnn.py:165-257]; expects (B,6,N), raises ValueError otherwise."""

    def __init__(self, num_classes=13, k=20, emb_dims=1024, dropout=0.5):
        super().__init__()
        self.k = k
        self.num_classes = num_classes
        self.conv1 = EdgeConv(3, 64, k)
        self.conv2 = EdgeConv(64, 64, k)
        self.conv3 = EdgeConv(64, 64, k)
        self.conv4 = EdgeConv(64, 128, k)
        self.color_conv = _pointwise(3, 64)
        self.conv5 = _pointwise(384, emb_dims)
        self.conv6 = _pointwise(emb_dims + 384, 512, dropout)
        self.conv7 = _pointwise(512, 256, dropout)
        self.conv8 = nn.Conv1d(256, num_classes, kernel_size=1)

    def prepare_geometry(self, x, stream=None):
        return _first_layer_graph(x[:, :3, :], self.k, stream)

    def forward(self, x, geometry=None, lengths=None):
        """lengths: as DGCNN.forward."""
        if x.size(1) != 6:
            raise ValueError("DGCNNWithColor expects 6-channel input (xyz + rgb)")
        x1 = self.conv1(x[:, :3, :], _nbr=_graph_from(geometry, x.shape[2]), lengths=lengths)
        x2 = self.conv2(x1, lengths=lengths)
        x3 = self.conv3(x2, lengths=lengths)
        x4 = self.conv4(x3, lengths=lengths)
        # the head runs point-major: (B,N,C) rows, one GEMM per layer, logits come out as (B,N,classes)
        color = _run_pointwise(self.color_conv, x[:, 3:6, :].permute(0, 2, 1))
        r_cat = torch.cat([t.permute(0, 2, 1) for t in (x1, x2, x3, x4)] + [color], dim=2)
        r5 = _run_pointwise(self.conv5, r_cat)
        r7 = _run_pointwise(self.conv7, _run_pointwise_cat(self.conv6, r_cat, r5))
        logits = _run_pointwise(self.conv8, r7)
        return logits, r5.permute(0, 2, 1), None


def get_model(num_classes=13, use_color=True, **kwargs):
    """[dgcnn.py:260-273]"""
    return DGCNNWithColor(num_classes=num_classes, **kwargs) if use_color else DGCNN(num_classes=num_classes, **kwargs)


def get_loss():
    """[dgcnn.py:276-280]"""
    return nn.CrossEntropyLoss(ignore_index=-1)

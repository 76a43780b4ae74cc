"""Host-side mirror of the reference's Training/metrics.py on top of libpcnbr.

Same function names, arguments and return values as /root/reference/Training/metrics.py:3-146, but every metric is read
off ONE device-side confusion matrix (csrc/metrics.cu) instead of B*C (or B*C^2) Python iterations with an `.item()`
sync each.  `predictions` (B,N,C) are the class scores -- softmax or raw logits, only the argmax matters; `labels`
(B,N,C) one-hot uint8; `mask` (B,) the number of unpadded points of each cloud (the reference's name for `lengths`).
CUDA tensors only."""
from __future__ import annotations

import torch

from . import _lib
from .ops import _c, _stream

__all__ = ["confusion_matrix_device", "overall_accuracy", "update_accuracy", "confusion_matrix",
           "intersection_over_union", "update_intersection_over_union"]


def confusion_matrix_device(predictions: torch.Tensor, labels: torch.Tensor, mask: torch.Tensor | None,
                            out: torch.Tensor | None = None, unlabeled: torch.Tensor | None = None) -> torch.Tensor:
    """(C,C) int64 on the device, matrix[label, predicted]; accumulated into `out` when given (validation loops keep one
    matrix for the whole set and read it once).  No host synchronisation.  unlabeled (C,) int64, accumulated: predictions of
    the rows whose label row is all zero (class 0 for the matrix, no class for the IoU, see _iou_terms)."""
    if not predictions.is_cuda:
        raise RuntimeError("pcnbr: predictions must be a CUDA tensor (this build has no CPU fallback)")
    B, N, C = predictions.shape
    pred = _c(predictions.float())
    lab = _c(labels.to(device=pred.device, dtype=torch.uint8))
    lens = None if mask is None else _c(mask.to(device=pred.device, dtype=torch.int64))
    m = torch.zeros(C, C, dtype=torch.int64, device=pred.device) if out is None else out
    _lib.call("pcnbr_confusion_ex_f32", pred.data_ptr(), lab.data_ptr(), lens.data_ptr() if lens is not None else None,
              B, N, C, m.data_ptr(), unlabeled.data_ptr() if unlabeled is not None else None, _stream())
    return m


def _iou_terms(m: torch.Tensor, unl: torch.Tensor):
    """(intersections, unions) of Training/metrics.py:97-110,131-144 from the confusion matrix and the unlabeled-row counts:
    the reference tests `labels[..., c] == 1`, so a row without any label is in no class's label set -- it only enlarges
    the union of the class it was predicted as -- while the matrix files it under label 0."""
    inter = m.diagonal().clone()
    rows = m.sum(dim=1)
    inter[0] -= unl[0]
    rows[0] -= unl.sum()
    union = rows + m.sum(dim=0) - inter
    return inter, union


def update_accuracy(predictions, labels, mask):
    """-> (correct points, total points) of the batch   [metrics.py:28-50]."""
    m = confusion_matrix_device(predictions, labels, mask)
    both = torch.stack((m.diagonal().sum(), m.sum())).tolist()           # one D2H read
    return both[0], both[1]


def overall_accuracy(predictions, labels, mask) -> float:
    """[metrics.py:3-25]"""
    correct, total = update_accuracy(predictions, labels, mask)
    return correct / total


def confusion_matrix(predictions, labels, mask) -> torch.Tensor:
    """(C,C) int64 CPU tensor, rows = label, columns = prediction   [metrics.py:52-78]."""
    return confusion_matrix_device(predictions, labels, mask).cpu()


def update_intersection_over_union(predictions, labels, mask):
    """-> (intersections (C,), unions (C,)) float32 CPU tensors   [metrics.py:115-146]."""
    unl = torch.zeros(predictions.shape[-1], dtype=torch.int64, device=predictions.device)
    m = confusion_matrix_device(predictions, labels, mask, unlabeled=unl)
    inter, union = _iou_terms(m, unl)
    both = torch.stack((inter, union)).to(torch.float32).cpu()
    return both[0], both[1]


def intersection_over_union(predictions, labels, mask):
    """-> (mean IoU, per-class IoU (C,)) with the reference's eps = 1e-6   [metrics.py:81-112]."""
    inter, union = update_intersection_over_union(predictions, labels, mask)
    eps = 1e-6
    ious = ((inter.double() + eps) / (union.double() + eps)).to(torch.float32)
    return ious.mean().item(), ious

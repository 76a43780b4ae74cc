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
                            out: torch.Tensor | None = None) -> torch.Tensor:
    """(C,C) int64 on the device, matrix[label, predicted]; accumulated into `out` when given (validation loops keep one
    matrix for the whole set and read it once).  No host synchronisation."""
    if not predictions.is_cuda:
        raise RuntimeError("pcnbr: predictions must be a CUDA tensor (this build has no CPU fallback)")
    B, N, C = predictions.shape
    pred = _c(predictions.float())
    lab = _c(labels.to(device=pred.device, dtype=torch.uint8))
    lens = None if mask is None else _c(mask.to(device=pred.device, dtype=torch.int64))
    m = torch.zeros(C, C, dtype=torch.int64, device=pred.device) if out is None else out
    _lib.call("pcnbr_confusion_f32", pred.data_ptr(), lab.data_ptr(), lens.data_ptr() if lens is not None else None,
              B, N, C, m.data_ptr(), _stream())
    return m


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
    m = confusion_matrix_device(predictions, labels, mask)
    inter = m.diagonal()
    union = m.sum(dim=0) + m.sum(dim=1) - inter
    both = torch.stack((inter, union)).to(torch.float32).cpu()
    return both[0], both[1]


def intersection_over_union(predictions, labels, mask):
    """-> (mean IoU, per-class IoU (C,)) with the reference's eps = 1e-6   [metrics.py:81-112]."""
    inter, union = update_intersection_over_union(predictions, labels, mask)
    eps = 1e-6
    ious = ((inter.double() + eps) / (union.double() + eps)).to(torch.float32)
    return ious.mean().item(), ious

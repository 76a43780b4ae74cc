"""Generate tests/golden/metrics.pt by running the UNMODIFIED reference Training/metrics.py (imported from
/root/reference).  TEST INFRASTRUCTURE ONLY; run in the build container:

    python oracle/make_golden_metrics.py
"""
from __future__ import annotations

import os
import sys

import torch

REF = os.environ.get("PCNBR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from Training import metrics as RM                      # noqa: E402  (reference)
from oracle import ref_ops as O                          # noqa: E402


def main():
    g = torch.Generator().manual_seed(20261018)
    B, N, C = 3, 257, 13
    logits = torch.randn(B, N, C, generator=g)
    logits[0, 5, 2] = logits[0, 5, 7] = logits[0, 5].max() + 1.0          # a tie: torch.argmax takes the first
    pred = torch.softmax(logits, dim=-1)
    lab = torch.nn.functional.one_hot(torch.randint(0, C, (B, N), generator=g), C).to(torch.uint8)
    mask = torch.tensor([257, 100, 0], dtype=torch.int64)                 # full, padded and empty cloud
    cm = RM.confusion_matrix(pred, lab, mask)
    correct, total = RM.update_accuracy(pred, lab, mask)
    inter, union = RM.update_intersection_over_union(pred, lab, mask)
    miou, ious = RM.intersection_over_union(pred, lab, mask)
    acc = RM.overall_accuracy(pred, lab, mask)
    # the restatement agrees with the reference on the same input
    assert torch.equal(cm, O.metrics_confusion_matrix(pred, lab, mask))
    assert (correct, total) == O.metrics_update_accuracy(pred, lab, mask)
    oi, ou = O.metrics_update_iou(pred, lab, mask)
    assert torch.equal(inter, oi) and torch.equal(union, ou)
    torch.save(dict(pred=pred, labels=lab, mask=mask, confusion=cm, correct=correct, total=total, inter=inter, union=union,
                    miou=miou, ious=ious, acc=acc), os.path.join(ROOT, "tests", "golden", "metrics.pt"))
    print("wrote tests/golden/metrics.pt", cm.sum().item(), "points, acc", acc, "mIoU", miou)
    # ---- malformed label rows (all zero): the confusion matrix / accuracy file them under class 0 (labels.argmax, :20,72),
    # the IoU functions under no class (labels[..., c] == 1, :103,137)
    lab2 = lab.clone()
    lab2[0, ::7] = 0
    lab2[1, :40] = 0
    mask2 = torch.tensor([257, 100, 5], dtype=torch.int64)
    cm2 = RM.confusion_matrix(pred, lab2, mask2)
    c2, t2 = RM.update_accuracy(pred, lab2, mask2)
    i2, u2 = RM.update_intersection_over_union(pred, lab2, mask2)
    m2, ious2 = RM.intersection_over_union(pred, lab2, mask2)
    assert torch.equal(cm2, O.metrics_confusion_matrix(pred, lab2, mask2)) and (c2, t2) == O.metrics_update_accuracy(pred, lab2, mask2)
    oi, ou = O.metrics_update_iou(pred, lab2, mask2)
    assert torch.equal(i2, oi) and torch.equal(u2, ou)
    torch.save(dict(pred=pred, labels=lab2, mask=mask2, confusion=cm2, correct=c2, total=t2, inter=i2, union=u2, miou=m2, ious=ious2),
               os.path.join(ROOT, "tests", "golden", "metrics_unlabeled.pt"))
    print("wrote tests/golden/metrics_unlabeled.pt")


if __name__ == "__main__":
    main()

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


if __name__ == "__main__":
    main()

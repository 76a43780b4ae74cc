"""GPU: the device-side evaluation metrics (csrc/metrics.cu) against the reference's golden outputs
and the oracle restatement of Training/metrics.py, and the north-star mIoU bar (unchanged to within 0.1) for the
B200 PointNet++ against the oracle model with the same weights."""
import pytest
import torch

from oracle import ref_ops as O

pytestmark = pytest.mark.gpu


def test_metrics_golden(pkg, dev, golden):
    g = golden("metrics")
    M = pkg.metrics
    pred, lab, mask = g["pred"].to(dev), g["labels"].to(dev), g["mask"].to(dev)
    assert torch.equal(M.confusion_matrix(pred, lab, mask), g["confusion"])
    assert M.update_accuracy(pred, lab, mask) == (g["correct"], g["total"])
    assert M.overall_accuracy(pred, lab, mask) == g["acc"]
    inter, union = M.update_intersection_over_union(pred, lab, mask)
    assert torch.equal(inter, g["inter"]) and torch.equal(union, g["union"])
    miou, ious = M.intersection_over_union(pred, lab, mask)
    assert torch.equal(ious, g["ious"]) and abs(miou - g["miou"]) < 1e-7


@pytest.mark.parametrize("B,N,C", [(32, 4096, 13), (2, 24000, 14), (5, 333, 3), (1, 7, 64)])
def test_confusion_matrix_vs_oracle_and_accumulation(pkg, dev, B, N, C):
    g = torch.Generator().manual_seed(B * N + C)
    logits = torch.randn(B, N, C, generator=g).round(decimals=1)              # plenty of exact ties
    lab = torch.nn.functional.one_hot(torch.randint(0, C, (B, N), generator=g), C).to(torch.uint8)
    mask = torch.randint(0, N + 1, (B,), generator=g)
    mask[0] = N
    want = O.metrics_confusion_matrix(logits, lab, mask)
    acc = torch.zeros(C, C, dtype=torch.int64, device=dev)
    for _ in range(2):                                                         # accumulates into `out`
        pkg.metrics.confusion_matrix_device(logits.to(dev), lab.to(dev), mask.to(dev), out=acc)
    assert torch.equal(acc.cpu(), 2 * want)
    assert int(want.sum()) == int(mask.sum())
    full = pkg.metrics.confusion_matrix_device(logits.to(dev), lab.to(dev), None)
    assert torch.equal(full.cpu(), O.metrics_confusion_matrix(logits, lab, torch.full((B,), N)))


def test_cpu_tensors_raise(pkg):
    with pytest.raises(RuntimeError):
        pkg.metrics.confusion_matrix_device(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3, dtype=torch.uint8), None)


def test_pointnetpp_miou_matches_oracle_model_within_bar(pkg, dev):
    """North star: mIoU unchanged to within 0.1 (percentage points) between the reference path and ours on the same
    inputs and weights.  Eval-mode PointNet++ on S3DIS-shaped blocks: logits -> softmax -> Training/metrics.py IoU."""
    pts, lab, lens = O.s3dis_blocks(4, 4096, seed=11)
    torch.manual_seed(5)
    ref = O.PointNetpp(13, tie="canon").eval()
    net = pkg.PointNetpp(13)
    net.load_state_dict(ref.state_dict())
    net = net.to(dev).eval()
    st = torch.tensor([1, 2, 3, 4], dtype=torch.int32)
    for name in ("sa1", "sa2", "sa3", "sa4"):
        getattr(net, name).fps_start = st.to(dev)
        getattr(ref, name).fps_start = st
    with torch.no_grad():
        p_ref = torch.softmax(ref(pts), dim=-1)
        p_our = torch.softmax(net(pts.to(dev)), dim=-1)
    miou_ref, _ = O.metrics_iou(p_ref, lab, lens)
    miou_our, _ = pkg.metrics.intersection_over_union(p_our, lab.to(dev), lens.to(dev))
    agree = (p_ref.argmax(-1) == p_our.cpu().argmax(-1)).float().mean().item()
    assert abs(100.0 * miou_our - 100.0 * miou_ref) <= 0.1, (miou_our, miou_ref)
    assert agree > 0.999


@pytest.mark.parametrize("B,L,C", [(16, 4096, 13), (3, 1000, 14), (2, 77, 5)])
def test_masked_cross_entropy_value_and_gradient(pkg, dev, B, L, C):
    """train.masked_onehot_cross_entropy on CUDA = Training/train_model.py:15-57 (log_softmax, -sum(onehot*logp), position
    mask, masked mean) in one fused kernel with its gradient: against the float64 formula, padded and empty clouds."""
    g = torch.Generator().manual_seed(B * L + C)
    logits = torch.randn(B, L, C, generator=g) * 3.0
    lab = torch.nn.functional.one_hot(torch.randint(0, C, (B, L), generator=g), C).to(torch.uint8)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    lens[0], lens[-1] = L, 0
    x64 = logits.double().requires_grad_(True)
    logp = torch.log_softmax(x64, dim=-1)
    tok = -(lab.double() * logp).sum(-1)
    mask = (torch.arange(L).unsqueeze(0) < lens.unsqueeze(1)).double()
    want = (tok * mask).sum() / mask.sum()
    (want * 2.5).backward()
    xd = logits.to(dev).requires_grad_(True)
    launches0 = pkg._lib.launches
    got = pkg.train.masked_onehot_cross_entropy(xd, lab.to(dev), lens.to(dev))
    assert pkg._lib.launches == launches0 + 2
    (got * 2.5).backward()
    assert abs(got.item() - want.item()) <= 2e-6 * abs(want.item())
    err = (xd.grad.cpu().double() - x64.grad).abs().max().item()
    assert err <= 1e-6 * x64.grad.abs().max().item() + 1e-12
    # every point padded: loss 0, gradient 0 (train_model.py:53-54)
    z = pkg.train.masked_onehot_cross_entropy(xd.detach().requires_grad_(True), lab.to(dev), torch.zeros(B, dtype=torch.int64, device=dev))
    assert z.item() == 0.0

"""GPU: the drop-in modules and models (same parameters loaded) against the reference's golden
logits/gradients and against the torch-CPU oracle models: fp32 features, logits and gradients within
1e-4 relative (north star), at shapes up to the BASELINE configs."""
import pytest
import torch

from oracle import ref_ops as O

pytestmark = pytest.mark.gpu


def _close(a, b, rtol=1e-4, scale_atol=1e-4):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    atol = scale_atol * b.abs().max().item()
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f"max abs err {(a - b).abs().max().item():.3e} vs scale {b.abs().max().item():.3e}"


def test_edgeconv_golden(pkg, dev, golden):
    g = golden("edgeconv")
    torch.manual_seed(g["seed"])
    ec = pkg.dgcnn.EdgeConv(g["cin"], g["cout"], k=g["k"]).to(dev)
    _close(ec(g["x"].to(dev)), g["out"])


def test_set_abstraction_and_fp_golden(pkg, dev, golden):
    g = golden("set_abstraction")
    torch.manual_seed(g["seed"])
    sa = pkg.common.SetAbstraction(g["C"], g["r"], g["cin"], g["mlps"]).to(dev)
    sa.fps_start = g["start"].to(dev)
    c1, f1 = sa(g["coords"].to(dev), g["features"].to(dev))
    assert torch.equal(c1.cpu(), g["centroids"])
    _close(f1, g["out"])
    h = golden("feature_propagation")
    torch.manual_seed(h["seed"])
    fp = pkg.common.FeaturePropagation(h["cin"], h["mlps"]).to(dev)
    out = fp(h["coords_1"].to(dev), h["coords_2"].to(dev), h["features_1"].to(dev), h["features_2"].to(dev))
    _close(out, h["out"])


def test_pointnetpp_golden_logits_and_grads(pkg, dev, golden):
    g = golden("pointnetpp")
    torch.manual_seed(g["seed"])
    net = pkg.PointNetpp(13)
    net.drop.p = 0.0
    net = net.to(dev)
    for sa, st in zip((net.sa1, net.sa2, net.sa3, net.sa4), g["fps_starts"]):
        sa.fps_start = st.to(dev)
    logits = net(g["x"].to(dev))
    _close(logits, g["logits"])
    (logits * g["loss_weight"].to(dev)).sum().backward()
    params = dict(net.named_parameters())
    for k, v in g["grads"].items():
        _close(params[k].grad, v, rtol=1e-3, scale_atol=1e-3)


@pytest.mark.parametrize("name", ["dgcnn", "dgcnn_color"])
def test_dgcnn_golden_logits_and_grads(pkg, dev, golden, name):
    g = golden(name)
    cls = pkg.DGCNN if name == "dgcnn" else pkg.DGCNNWithColor
    torch.manual_seed(g["seed"])
    m = cls(num_classes=13, k=g["k"], emb_dims=g["emb_dims"], dropout=0.0).to(dev)
    logits, emb, third = m(g["x"].to(dev))
    assert third is None
    _close(logits, g["logits"])
    _close(emb, g["emb"])
    (logits * g["loss_weight"].to(dev)).sum().backward()
    params = dict(m.named_parameters())
    for k, v in g["grads"].items():
        _close(params[k].grad, v, rtol=1e-3, scale_atol=1e-3)


def _copy_model(ours, oracle_model):
    ours.load_state_dict(oracle_model.state_dict())


def test_pointnetpp_s3dis_block_vs_oracle_model(pkg, dev):
    """BASELINE config 0 shape: batch 2 x 4096 x 9, 13 classes, under-filled balls (canonical ties)."""
    pts, _, _ = O.s3dis_blocks(2, 4096, seed=0)
    torch.manual_seed(3)
    ref = O.PointNetpp(13, tie="canon")
    ref.drop.p = 0.0
    net = pkg.PointNetpp(13)
    net.drop.p = 0.0
    _copy_model(net, ref)
    net = net.to(dev)
    starts = [torch.tensor([1, 2], dtype=torch.int32)] * 4
    for a, b, st in zip((net.sa1, net.sa2, net.sa3, net.sa4), (ref.sa1, ref.sa2, ref.sa3, ref.sa4), starts):
        a.fps_start, b.fps_start = st.to(dev), st
    w = torch.randn(2, 4096, 13, generator=torch.Generator().manual_seed(1))
    lo = ref(pts)
    (lo * w).sum().backward()
    lg = net(pts.to(dev))
    (lg * w.to(dev)).sum().backward()
    _close(lg, lo)
    pr = dict(ref.named_parameters())
    for k, p in net.named_parameters():
        _close(p.grad, pr[k].grad, rtol=1e-3, scale_atol=1e-3)


def test_pointnext_vs_oracle_model(pkg, dev):
    pts, _, _ = O.s3dis_blocks(2, 2048, seed=5)
    torch.manual_seed(4)
    ref = O.PointNeXt(13, tie="canon")
    ref.drop.p = 0.0
    net = pkg.PointNeXt(13)
    net.drop.p = 0.0
    _copy_model(net, ref)
    net = net.to(dev)
    st = torch.tensor([0, 7], dtype=torch.int32)
    for name in ("sa1", "sa2", "sa3", "sa4"):
        getattr(net, name).fps_start, getattr(ref, name).fps_start = st.to(dev), st
    lo = ref(pts)
    lg = net(pts.to(dev))
    _close(lg, lo)


def test_dgcnn_color_full_width_vs_oracle_model(pkg, dev):
    """DGCNNWithColor k=20, emb 1024 on an S3DIS-shaped block (xyz with room offsets + rgb)."""
    pts, _, _ = O.s3dis_blocks(2, 1024, seed=7)
    x = pts[:, :, :6].transpose(1, 2).contiguous()
    torch.manual_seed(5)
    ref = O.DGCNNWithColor(13, k=20, dropout=0.0, tie="canon")
    net = pkg.DGCNNWithColor(13, k=20, dropout=0.0)
    _copy_model(net, ref)
    net = net.to(dev)
    w = torch.randn(2, 1024, 13, generator=torch.Generator().manual_seed(2))
    lo = ref(x)[0]
    (lo * w).sum().backward()
    lg = net(x.to(dev))[0]
    (lg * w.to(dev)).sum().backward()
    _close(lg, lo)
    pr = dict(ref.named_parameters())
    for k, p in net.named_parameters():
        _close(p.grad, pr[k].grad, rtol=1e-3, scale_atol=1e-3)

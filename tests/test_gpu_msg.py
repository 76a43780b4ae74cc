"""GPU: multi-radius ball query and the multi-scale-grouping set abstraction (BASELINE configs[2], "PointNet++ MSG").
The reference has no MSG class: MSG = several `group` calls on one centroid set, so the oracle is the
reference's ball query per scale (oracle/canon.c, bit-exact) and the composition of its own blocks (oracle/ref_ops.py)."""
import pytest
import torch

from oracle import canon
from oracle import ref_ops as O
from test_gpu_models import _as_good_as_reference, _close, _copy_model, _fp64_twin

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,M,radii,Ks", [
    (4096, 1024, [0.05, 0.1], [16, 32]),                 # PointNetppMSG levels
    (1024, 256, [0.1, 0.2], [16, 32]),
    (256, 64, [0.2, 0.4], [16, 32]),
    (64, 16, [0.4, 0.8], [16, 32]),
    (4096, 512, [0.1, 0.2, 0.4], [16, 32, 128]),        # three scales, the classic MSG classification set-up
    (1000, 100, [0.3, 0.1, 0.2], [16, 48, 32]),         # unordered radii, the SMALLEST radius wants the most points
    (700, 50, [0.2, 0.2], [8, 64]),                      # same ball twice, different K
    (500, 20, [0.5], [100]),                             # one scale = plain ball query
    (33, 5, [0.05, 0.1, 0.2, 5.0], [33, 1, 20, 33]),     # K = N, a ball that holds everything
])
@pytest.mark.parametrize("kind", ["s3dis", "filled"])
def test_multi_radius_ball_query_is_per_scale_ball_query(pkg, dev, N, M, radii, Ks, kind):
    if kind == "s3dis":                                   # under-filled balls: every table ends in canonical padding
        xyz = O.s3dis_blocks(2, N, seed=M)[0][:, :, :3].contiguous()
    else:                                                 # every r = 0.1 ball holds >= K points
        xyz = torch.rand(2, N, 3, generator=torch.Generator().manual_seed(N + M)) * 0.3
    cen = canon.fps(xyz, M, torch.zeros(2, dtype=torch.int32))[1]
    tables = pkg.ops.query_ball_point_multi(radii, Ks, xyz.to(dev), cen.to(dev))
    assert len(tables) == len(radii)
    for r, K, idx in zip(radii, Ks, tables):
        assert idx.shape == (2, M, K) and idx.dtype == torch.int32
        assert torch.equal(idx.cpu(), canon.ball_query(cen, xyz, r, K)), f"scale r={r} K={K}"
        assert torch.equal(idx, pkg.ops.query_ball_point(r, K, xyz.to(dev), cen.to(dev)))


def test_multi_radius_lattice_ties_and_duplicates(pkg, dev):
    g = torch.Generator().manual_seed(9)
    xyz = torch.randint(-20, 21, (2, 600, 3), generator=g).float() / 256      # exact arithmetic, many equal distances
    xyz[1, 400:] = xyz[1, :200]                                               # duplicated points
    q = xyz[:, :70].contiguous()
    radii, Ks = [0.05, 0.08, 0.2], [16, 32, 48]
    for r, K, idx in zip(radii, Ks, pkg.ops.query_ball_point_multi(radii, Ks, xyz.to(dev), q.to(dev))):
        assert torch.equal(idx.cpu(), canon.ball_query(q, xyz, r, K))


def test_multi_radius_errors(pkg, dev):
    xyz = torch.rand(1, 10, 3, device=dev)
    with pytest.raises(RuntimeError):
        pkg.ops.query_ball_point_multi([0.1, 0.2], [4, 11], xyz, xyz)                      # K > N: torch.topk's error
    with pytest.raises(ValueError):
        pkg.ops.query_ball_point_multi([0.1, 0.2], [4], xyz, xyz)
    with pytest.raises(RuntimeError):
        pkg.ops.query_ball_point_multi([0.1], [4], xyz.cpu(), xyz.cpu())                    # no CPU fallback


@pytest.mark.parametrize("grouping_norm", [False, True])
def test_set_abstraction_msg_vs_oracle(pkg, dev, grouping_norm):
    B, N, D = 2, 1024, 6
    pts = O.s3dis_blocks(B, N, seed=11)[0]
    xyz, feat = pts[:, :, :3].contiguous(), pts[:, :, 3:].contiguous() / 255.0
    torch.manual_seed(2)
    ref = O.SetAbstractionMSG(256, [0.1, 0.2], D + 3, [[16, 16, 32], [32, 32, 64]], [16, 32], grouping_norm=grouping_norm)
    net = pkg.common.SetAbstractionMSG(256, [0.1, 0.2], D + 3, [[16, 16, 32], [32, 32, 64]], [16, 32], grouping_norm=grouping_norm)
    assert set(net.state_dict().keys()) == set(ref.state_dict().keys())
    _copy_model(net, ref)
    net = net.to(dev)
    st = torch.tensor([3, 9], dtype=torch.int32)
    net.fps_start, ref.fps_start = st.to(dev), st
    f_ref = feat.clone().requires_grad_(True)
    f_dev = feat.to(dev).requires_grad_(True)
    c_ref, o_ref = ref(xyz, f_ref)
    c_dev, o_dev = net(xyz.to(dev), f_dev)
    assert torch.equal(c_dev.cpu(), c_ref) and o_dev.shape == (B, 256, 96)
    _close(o_dev, o_ref)
    w = torch.randn(o_ref.shape, generator=torch.Generator().manual_seed(5))
    (o_ref * w).sum().backward()
    (o_dev * w.to(dev)).sum().backward()
    _close(f_dev.grad, f_ref.grad, rtol=1e-3, scale_atol=2e-4)
    with pytest.raises(ValueError):
        pkg.common.SetAbstractionMSG(16, [0.1], 9, [[8], [8]], [4])


def test_pointnetpp_msg_vs_oracle_model(pkg, dev):
    """BASELINE configs[2] network at 2 x 4096 x 9: as close to the float64 evaluation of the oracle model as the
    oracle's own fp32 path (the bar of the SSG model test)."""
    pts, _, _ = O.s3dis_blocks(2, 4096, seed=0)
    torch.manual_seed(3)
    ref = O.PointNetppMSG(13, tie="canon")
    ref.drop.p = 0.0
    net = pkg.PointNetppMSG(13)
    net.drop.p = 0.0
    assert list(net.state_dict().keys()) == list(ref.state_dict().keys())
    _copy_model(net, ref)
    net = net.to(dev)
    ref64 = _fp64_twin(ref)
    st = torch.tensor([1, 2], dtype=torch.int32)
    for name in ("sa1", "sa2", "sa3", "sa4"):
        getattr(net, name).fps_start = st.to(dev)
        getattr(ref, name).fps_start = getattr(ref64, name).fps_start = st
    w = torch.randn(2, 4096, 13, generator=torch.Generator().manual_seed(1))
    lo = ref(pts)
    (lo * w).sum().backward()
    lo64 = ref64(pts.double())
    (lo64 * w.double()).sum().backward()
    lg = net(pts.to(dev))
    (lg * w.to(dev)).sum().backward()
    _as_good_as_reference(lg, lo, lo64, "logits")
    pr, pr64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
    for k, p in net.named_parameters():
        _as_good_as_reference(p.grad, pr[k].grad, pr64[k].grad, f"grad {k}")


def test_set_abstraction_msg_golden(pkg, dev, golden):
    """Against the composition of the unmodified reference's own functions (oracle/make_golden_msg.py): centroids and
    all three neighbour tables bit-exact, features to 1e-4."""
    g = golden("msg")
    xyz, feat = g["coords"].to(dev), g["features"].to(dev)
    tables = pkg.ops.query_ball_point_multi(g["radii"], g["Ks"], xyz, g["centroids"].to(dev))
    for got, want in zip(tables, g["tables"]):
        assert torch.equal(got.cpu(), want)
    torch.manual_seed(g["seed"])
    net = pkg.common.SetAbstractionMSG(g["C"], g["radii"], g["cin"], g["mlps"], g["Ks"]).to(dev)
    net.fps_start = g["start"].to(dev)
    cen, out = net(xyz, feat)
    assert torch.equal(cen.cpu(), g["centroids"])
    _close(out, g["out"])

"""GPU: the LENGTH-AWARE forms of the hot path (zero-padded variable-N evaluation batches,
/root/reference/data_processing/block_datasets.py:19-25, Training/training.py:80-133).

Contract: with `lengths`, cloud b consists of its first lengths[b] rows only; the real rows get exactly what the reference
computes when that cloud is passed ALONE, unpadded -- indices bit-exact, features to 1e-4 -- and the padding rows an
in-range filler.  Checked against (a) golden vectors made by running the unmodified reference once per cloud
(oracle/make_golden_lengths.py) and (b) oracle/canon.c per cloud on seeded S3DIS-shaped inputs, through every dispatch
branch (register / cluster / global-memory FPS, scan / cell-grid selection, tensor-core / CUDA-core feature kNN)."""
import pytest
import torch

from oracle import canon, ref_ops as O

pytestmark = pytest.mark.gpu


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _padded_blocks(B, N, lengths, seed):
    """S3DIS-shaped blocks, rows >= lengths[b] zeroed as collate_blocks does."""
    pts, _, _ = O.s3dis_blocks(B, N, seed=seed)
    for b, n in enumerate(lengths):
        pts[b, n:] = 0.0
    return pts


# ------------------------------------------------------------------ golden vectors from the unmodified reference

def test_golden_sample_group_interpolate(pkg, dev, golden):
    g = golden("lengths")
    L, xyz = g["lengths"], g["xyz"].to(dev)
    s = g["sample"]
    idx, cen = pkg.ops.farthest_point_sample(xyz, s["C"], s["start"].to(dev), return_coords=True, lengths=L)
    assert torch.equal(cen.cpu(), s["coords"])
    assert all(int(idx[b].max()) < int(L[b]) for b in range(len(L)))
    gr = g["group"]
    out = pkg.common.group(cen, xyz, gr["features"].to(dev), gr["r"], gr["K"], True, lengths=L)
    assert torch.equal(out.cpu(), gr["out"])
    it = g["interpolate"]
    up = pkg.common.interpolate(it["points"].to(dev), xyz, cen, lengths=L).cpu()
    for b, n in enumerate(L.tolist()):
        assert torch.allclose(up[b, :n], it["out"][b, :n], rtol=1e-4, atol=1e-6)
        assert torch.isfinite(up[b]).all()


@pytest.mark.parametrize("F", [3, 64])
def test_golden_feature_knn(pkg, dev, golden, F):
    g = golden("lengths")
    L, k = g["lengths"], g[f"knn_F{F}"]["k"]
    idx = pkg.dgcnn.knn(g[f"knn_F{F}"]["x"].to(dev), k, lengths=L).cpu()
    for b, n in enumerate(L.tolist()):
        assert torch.equal(idx[b, :n].int(), g[f"knn_F{F}"]["idx"][b, :n])
        assert torch.equal(idx[b, n:], torch.arange(k).expand(idx.shape[1] - n, k))          # filler rows, in range


def test_golden_pointnetpp_eval_logits(pkg, dev, golden):
    """Whole model, eval mode, one zero-padded batch with lengths == the reference on every cloud alone."""
    g = golden("lengths")
    p, L = g["pointnetpp_eval"], g["lengths"]
    torch.manual_seed(p["seed"])
    net = pkg.PointNetpp(13).to(dev).eval()
    for sa, st in zip((net.sa1, net.sa2, net.sa3, net.sa4), p["fps_starts"]):
        sa.fps_start = st.to(dev)
    with torch.no_grad():
        logits = net(p["x"].to(dev), lengths=L).cpu()
        plain = net(p["x"].to(dev)).cpu()                       # the reference's own behaviour: padding participates
    scale = p["logits"].abs().max().item()
    for b, n in enumerate(L.tolist()):
        err = (logits[b, :n] - p["logits"][b, :n]).abs().max().item()
        assert err <= 1e-4 * scale, f"cloud {b}: {err:.3e} vs scale {scale:.3e}"
    assert torch.isfinite(logits).all()
    # the padded clouds really differ without lengths (the zero rows are picked by FPS and fill the balls)
    assert (plain[1, :L[1]] - p["logits"][1, :L[1]]).abs().max().item() > 1e-3 * scale


# ------------------------------------------------------------------ oracle/canon.c per cloud, every dispatch branch

@pytest.mark.parametrize("N,C,lengths,branch", [
    (4096, 1024, [4096, 3000, 1500, 700], "fps_reg_kernel"),        # 700 < C: the reference re-picks point 0
    (1000, 64, [1000, 999, 33, 1], "fps_reg_kernel"),
    (24000, 512, [24000, 9000, 16001], "fps_cluster_kernel"),
    (70000, 48, [70000, 12345], "fps_cluster_kernel"),              # 16-CTA cluster
    (140000, 24, [140000, 70001], "fps_big_kernel"),
])
def test_fps_lengths_vs_oracle(pkg, dev, N, C, lengths, branch):
    B = len(lengths)
    g = _gen(N + C)
    xyz = torch.rand(B, N, 3, generator=g) * torch.tensor([1.0, 1.0, 3.0]) + 5.0
    for b, n in enumerate(lengths):
        xyz[b, n:] = 0.0
    start = torch.randint(0, N, (B,), generator=g, dtype=torch.int32)
    pkg._lib.prof_enable(True)
    pkg._lib.prof_collect()
    idx, coords = pkg.ops.farthest_point_sample(xyz.to(dev), C, start.to(dev), return_coords=True, lengths=lengths)
    ran = pkg._lib.prof_collect()
    pkg._lib.prof_enable(False)
    assert any(k.startswith(branch) for k in ran), f"expected {branch}, ran {sorted(ran)}"
    for b, n in enumerate(lengths):
        o_idx, o_coords = canon.fps(xyz[b:b + 1, :n].contiguous(), C, start[b:b + 1] % n)
        assert torch.equal(idx[b:b + 1].cpu(), o_idx), f"cloud {b}"
        assert torch.equal(coords[b:b + 1].cpu(), o_coords), f"cloud {b}"


@pytest.mark.parametrize("N,M,lengths", [(1024, 256, [1024, 700, 333, 40]),          # M x N scan
                                         (4096, 1024, [4096, 3000, 2047, 64])])      # cell grid
@pytest.mark.parametrize("r,K", [(0.1, 32), (0.2, 16)])
def test_ball_query_lengths_vs_oracle(pkg, dev, N, M, lengths, r, K):
    B = len(lengths)
    pts = _padded_blocks(B, N, lengths, seed=N + K)
    xyz = pts[:, :, :3].contiguous()
    cen = torch.stack([xyz[b, torch.randperm(n, generator=_gen(b))[:M].repeat((M + n - 1) // n)[:M]] for b, n in enumerate(lengths)])
    idx = pkg.ops.query_ball_point(r, K, xyz.to(dev), cen.to(dev), lengths=lengths).cpu()
    for b, n in enumerate(lengths):
        assert torch.equal(idx[b:b + 1], canon.ball_query(cen[b:b + 1], xyz[b:b + 1, :n].contiguous(), r, K)), f"cloud {b}"
    # query lengths: rows behind query_lengths[b] are the filler
    ql = [M, M // 2, 1, M - 1]
    idx2 = pkg.ops.query_ball_point(r, K, xyz.to(dev), cen.to(dev), lengths=lengths, query_lengths=ql).cpu()
    for b, m in enumerate(ql):
        assert torch.equal(idx2[b, :m], idx[b, :m])
        assert torch.equal(idx2[b, m:], torch.arange(K, dtype=torch.int32).expand(M - m, K))


def test_ball_query_multi_lengths_vs_oracle(pkg, dev):
    lengths = [4096, 2500, 100]
    B, N, M = 3, 4096, 512
    xyz = _padded_blocks(B, N, lengths, seed=5)[:, :, :3].contiguous()
    cen = torch.stack([xyz[b, torch.arange(M) % n] for b, n in enumerate(lengths)])
    outs = pkg.ops.query_ball_point_multi([0.05, 0.1], [16, 32], xyz.to(dev), cen.to(dev), lengths=lengths)
    for (r, K), idx in zip(((0.05, 16), (0.1, 32)), outs):
        for b, n in enumerate(lengths):
            assert torch.equal(idx[b:b + 1].cpu(), canon.ball_query(cen[b:b + 1], xyz[b:b + 1, :n].contiguous(), r, K))


@pytest.mark.parametrize("N,M,lengths", [(1024, 256, [1024, 600, 3]),               # scan
                                         (4096, 1024, [4096, 2500, 5])])             # cell grid
def test_three_nn_lengths_vs_oracle(pkg, dev, N, M, lengths):
    """interpolate(): the fine cloud (queries) is the padded one; and the symmetric case of padded sources."""
    B = len(lengths)
    fine = _padded_blocks(B, N, lengths, seed=N)[:, :, :3].contiguous()
    coarse = torch.stack([fine[b, torch.arange(M) % n] for b, n in enumerate(lengths)])
    idx, d2 = pkg.ops.knn_points(fine.to(dev), coarse.to(dev), 3, query_lengths=lengths)
    for b, n in enumerate(lengths):
        o_idx, o_d2 = canon.knn_direct(fine[b:b + 1, :n].contiguous(), coarse[b:b + 1], 3)
        assert torch.equal(idx[b:b + 1, :n].cpu(), o_idx) and torch.equal(d2[b:b + 1, :n].cpu(), o_d2)
        assert torch.equal(idx[b, n:].cpu(), torch.arange(3, dtype=torch.int32).expand(N - n, 3))
    # padded SOURCES (queries all real): 3 nearest among the first lengths[b] sources
    idx, d2 = pkg.ops.knn_points(coarse.to(dev), fine.to(dev), 3, src_lengths=lengths)
    for b, n in enumerate(lengths):
        o_idx, o_d2 = canon.knn_direct(coarse[b:b + 1], fine[b:b + 1, :n].contiguous(), 3)
        assert torch.equal(idx[b:b + 1].cpu(), o_idx) and torch.equal(d2[b:b + 1].cpu(), o_d2)


@pytest.mark.parametrize("F,N,lengths", [
    (3, 4096, [4096, 3001, 2048, 300]),       # tensor cores: skipped units, a partial last column tile, N % 32 != 0
    (64, 4096, [4096, 1337, 129]),
    (64, 1024, [1000, 1024, 20]),             # k == shortest cloud
    (64, 200, [200, 77]),                     # N < 256: CUDA-core kernel by dispatch
    (100, 300, [300, 250, 21]),               # F > 64: CUDA-core kernel by dispatch
])
def test_feature_knn_lengths_vs_oracle(pkg, dev, F, N, lengths):
    B, k = len(lengths), 20
    x = torch.randn(B, F, N, generator=_gen(F * N))
    if F == 3:
        x = x + torch.tensor([12.0, 7.0, 1.5])[None, :, None]          # S3DIS-like offsets: the bounds must stay valid
    for b, n in enumerate(lengths):
        x[b, :, n:] = 0.0
    for layout in ("channel_major", "point_major"):
        xin = x.to(dev) if layout == "channel_major" else x.to(dev).transpose(1, 2).contiguous().transpose(1, 2)
        idx = pkg.ops.knn_graph(xin, k, lengths=lengths).cpu()
        for b, n in enumerate(lengths):
            o = canon.knn_expand(x[b:b + 1, :, :n].contiguous(), k)[0]
            assert torch.equal(idx[b:b + 1, :n], o), f"{layout} cloud {b}"
            assert torch.equal(idx[b, n:], torch.arange(k, dtype=torch.int32).expand(N - n, k))


def test_lengths_shorter_than_k_raise_like_topk(pkg, dev):
    xyz = torch.rand(2, 256, 3).to(dev)
    with pytest.raises(RuntimeError, match="out of range"):
        pkg.ops.query_ball_point(0.1, 32, xyz, xyz[:, :8].contiguous(), lengths=[256, 31])
    with pytest.raises(RuntimeError, match="out of range"):
        pkg.ops.knn_graph(torch.rand(2, 3, 256).to(dev), 20, lengths=[256, 19])
    with pytest.raises(ValueError):
        pkg.ops.farthest_point_sample(xyz, 8, lengths=[256])


def test_dgcnn_eval_lengths_equals_each_cloud_alone(pkg, dev):
    """DGCNNWithColor, eval mode: one zero-padded batch with lengths == each cloud evaluated alone (own path, B = 1, the
    unpadded cloud) -- the per-cloud graph, and with it every logit of the real rows, does not see the padding."""
    lengths = [1024, 700, 300]
    pts = _padded_blocks(3, 1024, lengths, seed=9)
    x = pts[:, :, :6].transpose(1, 2).contiguous()
    torch.manual_seed(5)
    net = pkg.DGCNNWithColor(13, k=20, emb_dims=64).to(dev).eval()
    with torch.no_grad():
        batch = net(x.to(dev), lengths=lengths)[0].cpu()
        for b, n in enumerate(lengths):
            alone = net(x[b:b + 1, :, :n].contiguous().to(dev), lengths=[n])[0].cpu()
            scale = alone.abs().max().item()
            err = (batch[b, :n] - alone[0]).abs().amax(dim=1)
            # a feature-space graph is discontinuous: a last-bit difference upstream may swap a 20th / 21st neighbour of a
            # row; all but a handful of rows must agree to 1e-4, none may be far off
            assert (err > 1e-4 * scale).float().mean().item() <= 0.005 and err.max().item() <= 0.05 * scale, f"cloud {b}"
    assert torch.isfinite(batch).all()


def test_pointnext_eval_lengths_equals_each_cloud_alone(pkg, dev):
    lengths = [2048, 1200]
    pts = _padded_blocks(2, 2048, lengths, seed=11)
    torch.manual_seed(6)
    net = pkg.PointNeXt(13).to(dev).eval()
    st = torch.tensor([3, 5], dtype=torch.int32, device=dev)
    with torch.no_grad():
        for name in ("sa1", "sa2", "sa3", "sa4"):
            getattr(net, name).fps_start = st
        batch = net(pts.to(dev), lengths=lengths).cpu()
        for b, n in enumerate(lengths):
            for name in ("sa1", "sa2", "sa3", "sa4"):
                getattr(net, name).fps_start = st[b:b + 1]
            alone = net(pts[b:b + 1, :n].contiguous().to(dev)).cpu()
            scale = alone.abs().max().item()
            assert (batch[b, :n] - alone[0]).abs().max().item() <= 1e-4 * scale, f"cloud {b}"

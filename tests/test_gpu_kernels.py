"""GPU: every libpcnbr kernel (through the host layer -> ctypes -> C ABI) against the CPU oracle on the
same seeded inputs and against the reference's golden vectors.  Indices are compared bit-exactly;
fp32 values bit-exactly where the kernel follows the reference's rounding sequence, else to 1e-4
relative (the north-star tolerance)."""
import pytest
import torch

from oracle import canon, ref_ops as O

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-5


def _gen(seed):
    return torch.Generator().manual_seed(seed)


# --------------------------------------------------------------------------- K1 FPS

def test_fps_golden(pkg, dev, golden):
    g = golden("fps")
    out = pkg.common.sample(g["xyz"].to(dev), g["C"], g["start"].to(dev))
    assert torch.equal(out.cpu(), g["coords"])


@pytest.mark.parametrize("B,N,C", [(2, 64, 16), (3, 256, 64), (2, 300, 77), (2, 1024, 256), (2, 2000, 128),
                                   (2, 4096, 1024), (1, 8192, 64), (2, 9000, 40), (1, 24000, 32)])
def test_fps_vs_oracle(pkg, dev, B, N, C):
    pts, _, _ = O.s3dis_blocks(B, N, seed=N)
    xyz = pts[:, :, :3].contiguous()
    start = torch.randint(0, N, (B,), generator=_gen(C), dtype=torch.int32)
    idx, coords = pkg.ops.farthest_point_sample(xyz.to(dev), C, start.to(dev), return_coords=True)
    o_idx, o_coords = canon.fps(xyz, C, start)
    assert torch.equal(idx.cpu(), o_idx)
    assert torch.equal(coords.cpu(), o_coords)


def test_fps_duplicates_padding_and_more_picks_than_points(pkg, dev):
    pts, _, _ = O.s3dis_blocks(2, 500, seed=2)
    xyz = pts[:, :, :3].contiguous()
    xyz[0, 300:] = 0.0                       # zero padding (block_datasets.py:19-25)
    xyz[1, 250:] = xyz[1, :250]              # exact duplicates (block_datasets.py:122-125)
    start = torch.tensor([499, 0], dtype=torch.int32)
    for C in (100, 600):                     # C > N repeats picks, like the reference loop
        idx = pkg.ops.farthest_point_sample(xyz.to(dev), C, start.to(dev))
        assert torch.equal(idx.cpu(), canon.fps(xyz, C, start)[0])


def test_fps_default_start_draw_matches_reference_rng_use(pkg, dev):
    """common.py:22 draws torch.randint(0, N, (B,), dtype=torch.int, device=coords.device)."""
    xyz = torch.rand(3, 128, 3, generator=_gen(0)).to(dev)
    torch.manual_seed(5)
    a = pkg.ops.farthest_point_sample(xyz, 8)
    torch.manual_seed(5)
    start = torch.randint(0, 128, (3,), dtype=torch.int, device=dev)
    assert torch.equal(a[:, 0], start)
    assert torch.equal(a, pkg.ops.farthest_point_sample(xyz, 8, start))


# --------------------------------------------------------------------------- K2 ball query / K5 group

@pytest.mark.parametrize("name", ["group_norm0", "group_norm1"])
def test_group_golden(pkg, dev, golden, name):
    g = golden(name)
    out = pkg.common.group(g["centroids"].to(dev), g["coords"].to(dev), g["features"].to(dev), g["r"], g["K"], g["normalize"])
    assert torch.equal(out.cpu(), g["out"])


@pytest.mark.parametrize("N,M,K,r", [(4096, 1024, 32, 0.1), (1024, 256, 32, 0.2), (256, 64, 32, 0.4), (64, 16, 32, 0.8),
                                     (1000, 100, 16, 0.15), (700, 50, 64, 0.3), (500, 20, 100, 0.5), (33, 5, 33, 0.2)])
def test_ball_query_underfilled_vs_oracle(pkg, dev, N, M, K, r):
    pts, _, _ = O.s3dis_blocks(2, N, seed=M)
    xyz = pts[:, :, :3].contiguous()
    cen = canon.fps(xyz, M, torch.zeros(2, dtype=torch.int32))[1]
    idx = pkg.ops.query_ball_point(r, K, xyz.to(dev), cen.to(dev))
    assert torch.equal(idx.cpu(), canon.ball_query(cen, xyz, r, K))


def test_ball_query_lattice_ties_and_duplicates(pkg, dev):
    xyz = torch.randint(-20, 21, (2, 600, 3), generator=_gen(9)).float() / 256
    xyz[1, 400:] = xyz[1, :200]
    q = xyz[:, :70].contiguous()
    for r, K in ((0.05, 16), (0.08, 32), (0.2, 48)):
        idx = pkg.ops.query_ball_point(r, K, xyz.to(dev), q.to(dev))
        assert torch.equal(idx.cpu(), canon.ball_query(q, xyz, r, K))


def test_ball_query_radius_threshold_is_fp32_of_double_square(pkg, dev):
    """d2 <= fp32(double(r)**2): 0.1 -> 0.0099999998, not 0.1f*0.1f."""
    import numpy as np
    t = np.float32(0.1 ** 2)
    up = np.nextafter(t, np.float32(1))
    # points on the x axis at distance sqrt(t) and sqrt(up) cannot be hit exactly; use the oracle instead
    xyz = torch.zeros(1, 64, 3)
    xyz[0, :, 0] = torch.linspace(0.0999, 0.1001, 64)
    q = torch.zeros(1, 1, 3)
    idx = pkg.ops.query_ball_point(0.1, 64, xyz.to(dev), q.to(dev))
    assert torch.equal(idx.cpu(), canon.ball_query(q, xyz, 0.1, 64))
    assert float(up) > float(t)


@pytest.mark.parametrize("D", [0, 1, 6, 29, 32, 64, 131, 256])
def test_group_values_and_backward(pkg, dev, D):
    B, N, M, K = 2, 512, 64, 32
    pts, _, _ = O.s3dis_blocks(B, N, seed=D)
    xyz = pts[:, :, :3].contiguous()
    feat = torch.randn(B, N, D, generator=_gen(D))
    cen = canon.fps(xyz, M, torch.zeros(B, dtype=torch.int32))[1]
    idx = canon.ball_query(cen, xyz, 0.2, K)
    fd = feat.to(dev).requires_grad_(D > 0)
    nbr = pkg.ops.NeighborIndex(idx.to(dev), N)
    out = pkg.ops.group_points(xyz.to(dev), fd, cen.to(dev), nbr, 0.2)
    assert torch.equal(out.detach().cpu(), canon.group(cen, xyz, feat, idx, 0.2, True))
    if D == 0:
        return
    w = torch.randn(out.shape, generator=_gen(1))
    (out * w.to(dev)).sum().backward()
    fr = feat.clone().requires_grad_(True)
    (O.group(cen, xyz, fr, 0.2, K, True, idx=idx.long()) * w).sum().backward()
    assert torch.allclose(fd.grad.cpu(), fr.grad, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("D", [6, 64, 128, 29])
def test_group_padded_rows_equal_reference_layout(pkg, dev, D):
    """pad4=True (16-byte row pitch for the tensor-core GEMM that follows): the first 3+D columns are the reference
    tensor bit for bit, the pad columns are zero, and the feature gradient is unchanged."""
    B, N, M, K = 2, 400, 50, 16
    pts, _, _ = O.s3dis_blocks(B, N, seed=D)
    xyz = pts[:, :, :3].contiguous().to(dev)
    feat = torch.randn(B, N, D, generator=_gen(D)).to(dev)
    cen = xyz[:, :M].contiguous()
    nbr = pkg.ops.NeighborIndex(pkg.ops.query_ball_point(0.3, K, xyz, cen), N)
    f1, f2 = feat.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    ref = pkg.ops.group_points(xyz, f1, cen, nbr, 0.3)
    pad = pkg.ops.group_points(xyz, f2, cen, nbr, 0.3, pad4=True)
    W = 3 + D
    assert pad.shape[-1] == (32 if W <= 32 else (W + 3) // 4 * 4) and pad.is_contiguous()
    assert torch.equal(pad[..., :W], ref) and bool((pad[..., W:] == 0).all())
    w = torch.randn(pad.shape, generator=_gen(2)).to(dev)
    (ref * w[..., :W]).sum().backward()
    (pad * w).sum().backward()
    assert torch.equal(f1.grad, f2.grad)


def test_csr_is_sorted_inverse(pkg, dev):
    B, N, M, K = 2, 300, 150, 32
    idx = torch.randint(0, N, (B, M, K), generator=_gen(3), dtype=torch.int32)
    idx[0, :, :8] = torch.arange(8, dtype=torch.int32)          # heavy hitters: 150 entries each
    idx[1] = 7                                                   # one segment holds everything
    offsets, perm = pkg.ops.NeighborIndex(idx.to(dev), N).csr()
    offsets, perm = offsets.cpu(), perm.cpu()
    flat = idx.view(B, -1)
    for b in range(B):
        order = torch.sort(flat[b].long(), stable=True).indices.int()    # ascending (source, position)
        assert torch.equal(perm[b], order)
        counts = torch.bincount(flat[b].long(), minlength=N)
        assert torch.equal(offsets[b, 1:].long(), torch.cumsum(counts, 0)) and offsets[b, 0] == 0


def _check_csr(idx, N, offsets, perm):
    B = idx.shape[0]
    flat = idx.reshape(B, -1)
    for b in range(B):
        order = torch.sort(flat[b].long(), stable=True).indices.int()    # ascending (source, position)
        assert torch.equal(perm[b], order)
        counts = torch.bincount(flat[b].long(), minlength=N)
        assert torch.equal(offsets[b, 1:].long(), torch.cumsum(counts, 0)) and offsets[b, 0] == 0


@pytest.mark.parametrize("kind,B,N,M,K", [("knn", 2, 4096, 4096, 20), ("ball", 3, 4096, 1024, 32), ("nn3", 2, 1024, 4096, 3),
                                          ("knn", 2, 1000, 1000, 33), ("ball", 2, 777, 45, 16), ("rand", 2, 50, 1000, 7),
                                          ("knn", 1, 12345, 12345, 16)])
def test_csr_bitmap_transposition_all_table_kinds(pkg, dev, kind, B, N, M, K):
    """pcnbr_csr_build_rows (mark / scan / emit, csrc/csr.cu) == stable sort of the flattened table, for distinct-row
    tables (kNN, 3-NN), hub-heavy padded ball tables and caller-made tables with repeated sources inside a row; the flat
    entry point pcnbr_csr_build must give the same inverse."""
    g = _gen(N + M + K)
    if kind == "knn" or kind == "nn3":           # distinct sources per row, a few popular hubs
        score = torch.rand(B, M, N, generator=g)
        score[:, :, : max(1, N // 100)] += 0.5
        idx = score.topk(K, dim=2).indices.int()
    elif kind == "ball":                           # under-filled balls: a handful of in-ball points, then 0,1,2,... padding
        idx = torch.empty(B, M, K, dtype=torch.int32)
        for b in range(B):
            for m in range(M):
                c = int(torch.randint(1, K // 2, (1,), generator=g))
                inball = torch.randperm(N - K, generator=g)[:c] + K
                idx[b, m] = torch.cat((inball, torch.arange(K - c))).int()
    else:
        idx = torch.randint(0, N, (B, M, K), generator=g, dtype=torch.int32)
    offsets, perm = pkg.ops.NeighborIndex(idx.to(dev), N).csr()
    _check_csr(idx, N, offsets.cpu(), perm.cpu())
    # the flat entry point (no row structure given)
    E = M * K
    off2 = torch.empty(B, N + 1, dtype=torch.int32, device=dev)
    perm2 = torch.empty(B, E, dtype=torch.int32, device=dev)
    nb = pkg._lib.size("pcnbr_csr_ws_bytes", B, E, N)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    pkg._lib.call("pcnbr_csr_build", idx.to(dev).data_ptr(), B, E, N, off2.data_ptr(), perm2.data_ptr(), ws.data_ptr(), nb,
                  torch.cuda.current_stream().cuda_stream)
    assert torch.equal(off2, offsets) and torch.equal(perm2, perm)


def test_group_backward_is_deterministic(pkg, dev):
    B, N, M, K, D = 2, 512, 128, 32, 64
    pts, _, _ = O.s3dis_blocks(B, N, seed=3)
    xyz = pts[:, :, :3].contiguous().to(dev)
    cen = pkg.common.sample(xyz, M, torch.zeros(B, dtype=torch.int32, device=dev))
    feat = torch.randn(B, N, D, device=dev)
    grads = []
    for _ in range(3):
        f = feat.clone().requires_grad_(True)
        pkg.common.group(cen, xyz, f, 0.1, K).square().sum().backward()
        grads.append(f.grad.clone())
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])


# --------------------------------------------------------------------------- index_points / square_distance (north-star names)

@pytest.mark.parametrize("N,D,shape", [(1024, 6, (256, 32)), (4096, 64, (1024,)), (333, 130, (77, 5)), (64, 3, (16, 32)), (50, 1, (7,))])
def test_index_points_is_the_reference_gather_with_scatter_add_backward(pkg, dev, N, D, shape):
    """points[batch_indices, indices] of common.py:64-65,117: values bit-exact, backward = index_put_(accumulate=True)."""
    B = 3
    g = _gen(N + D)
    pts = torch.randn(B, N, D, generator=g)
    idx = torch.randint(0, N, (B, *shape), generator=g)
    idx[0].view(-1)[: min(8, idx[0].numel())] = 0                                        # a hub
    pd = pts.to(dev).requires_grad_(True)
    for ix in (idx.to(dev), idx.to(dev).int()):
        out = pkg.ops.index_points(pd, ix)
        want = pts[torch.arange(B).view(B, *([1] * len(shape))), idx]
        assert out.shape == want.shape and torch.equal(out.detach().cpu(), want)
    w = torch.randn(out.shape, generator=g)
    (out * w.to(dev)).sum().backward()
    pr = pts.clone().requires_grad_(True)
    (pr[torch.arange(B).view(B, *([1] * len(shape))), idx] * w).sum().backward()
    assert torch.allclose(pd.grad.cpu(), pr.grad, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("N,M", [(4096, 1024), (300, 77), (1, 1), (257, 513)])
def test_square_distance_matches_the_reference_expression(pkg, dev, N, M):
    pts, _, _ = O.s3dis_blocks(2, max(N, M), seed=N + M)
    src, dst = pts[:, :N, :3].contiguous(), pts[:, :M, :3].flip(1).contiguous()
    want = ((dst.unsqueeze(1) - src.unsqueeze(2)) ** 2).sum(dim=-1)                      # common.py:54-56
    assert torch.equal(pkg.ops.square_distance(src.to(dev), dst.to(dev)).cpu(), want)
    with pytest.raises(ValueError):
        pkg.ops.square_distance(src.to(dev), dst[:1].to(dev))


# --------------------------------------------------------------------------- K6 max-pool

@pytest.mark.parametrize("shape", [(2, 5, 7, 11), (2, 64, 32, 64), (3, 16, 32, 513), (1, 10, 1, 4)])
def test_reduce_max_contiguous_and_permuted(pkg, dev, shape):
    x = torch.randn(*shape, generator=_gen(0))
    x[0, 0, :, 0] = 1.5                                          # ties: first maximum wins
    for make in (lambda t: t, lambda t: t.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)):
        xd = make(x.to(dev)).requires_grad_(True)
        out = pkg.common.reduce(xd, "max")
        assert torch.equal(out.detach().cpu(), O.reduce(x, "max"))
        w = torch.randn(out.shape, generator=_gen(1))
        (out * w.to(dev)).sum().backward()
        xr = x.clone().requires_grad_(True)
        (O.reduce(xr, "max") * w).sum().backward()
        assert torch.equal(xd.grad.cpu(), xr.grad)


def test_reduce_golden_avg_and_errors(pkg, dev, golden):
    g = golden("reduce")
    assert torch.equal(pkg.common.reduce(g["x"].to(dev), "max").cpu(), g["max"])
    assert torch.allclose(pkg.common.reduce(g["x"].to(dev), "avg").cpu(), g["avg"], rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        pkg.common.reduce(g["x"].to(dev), "min")


@pytest.mark.parametrize("channels_last", [False, True])
def test_edgeconv_style_max_over_last_dim(pkg, dev, channels_last):
    x = torch.randn(2, 24, 50, 20, generator=_gen(4))
    xd = x.to(dev)
    if channels_last:
        xd = xd.contiguous(memory_format=torch.channels_last)
    xd.requires_grad_(True)
    out = pkg.ops.max_pool_neighbors(xd, -1)
    assert torch.equal(out.detach().cpu(), x.max(dim=-1)[0])
    w = torch.randn(out.shape, generator=_gen(5))
    (out * w.to(dev)).sum().backward()
    xr = x.clone().requires_grad_(True)
    (xr.max(dim=-1)[0] * w).sum().backward()
    assert torch.equal(xd.grad.cpu(), xr.grad)


# --------------------------------------------------------------------------- K3 / K8 interpolate

def test_interpolate_golden(pkg, dev, golden):
    g = golden("interpolate")
    out = pkg.common.interpolate(g["points"].to(dev), g["coords_1"].to(dev), g["coords_2"].to(dev))
    assert torch.equal(out.cpu(), g["out"])


@pytest.mark.parametrize("N,M,D,k", [(4096, 1024, 128, 3), (1024, 256, 256, 3), (64, 16, 512, 3), (333, 77, 50, 5), (100, 8, 7, 8),
                                     (335, 77, 64, 3), (1001, 40, 132, 3), (50, 9, 6, 3)])
def test_interpolate_vs_oracle_with_backward(pkg, dev, N, M, D, k):
    pts, _, _ = O.s3dis_blocks(2, N, seed=N + M)
    fine = pts[:, :, :3].contiguous()
    coarse = canon.fps(fine, M, torch.zeros(2, dtype=torch.int32))[1]      # coarse points are fine points: d2 = 0 rows
    feats = torch.randn(2, M, D, generator=_gen(D))
    idx, d2 = pkg.ops.knn_points(fine.to(dev), coarse.to(dev), k)
    o_idx, o_d2 = canon.knn_direct(fine, coarse, k)
    assert torch.equal(idx.cpu(), o_idx) and torch.equal(d2.cpu(), o_d2)
    fd = feats.to(dev).requires_grad_(True)
    out = pkg.common.interpolate(fd, fine.to(dev), coarse.to(dev), k)
    assert torch.equal(out.detach().cpu(), canon.interp(feats, o_idx, o_d2))
    w = torch.randn(out.shape, generator=_gen(2))
    (out * w.to(dev)).sum().backward()
    fr = feats.clone().requires_grad_(True)
    (O.interpolate(fr, fine, coarse, k) * w).sum().backward()
    assert torch.allclose(fd.grad.cpu(), fr.grad, rtol=RTOL, atol=ATOL)


def test_interpolate_division_is_correctly_rounded_on_adversarial_values(pkg, dev):
    """The k = 3 kernel divides by a per-point reciprocal (Markstein sequence, csrc/interp.cu) instead of issuing
    N*k*D div.rn: its quotients must stay the correctly rounded (f * w) / norm of common.py:122 for EVERY float -- random
    bit patterns (denormals, huge values, signed zeros), divisors with an all-ones significand, tiny and huge norms."""
    B, N, M, D = 2, 3000, 64, 128
    g = _gen(99)
    bits = torch.randint(-2**31, 2**31 - 1, (B, M, D), generator=g, dtype=torch.int64).to(torch.int32)
    feats = bits.view(torch.float32).clone()
    feats[~torch.isfinite(feats)] = 0.0
    feats[:, :8] = torch.randn(B, 8, D, generator=g)                       # ordinary rows as well
    feats[:, 8:12] = torch.randn(B, 4, D, generator=g) * 1e-38             # denormal products
    idx = torch.randint(0, M, (B, N, 3), generator=g, dtype=torch.int32)
    d2 = torch.rand(B, N, 3, generator=g) * (10.0 ** torch.randint(-12, 5, (B, N, 3), generator=g).float())
    d2[:, :50] = 0.0                                                       # w = 1e9: the largest norms
    d2[:, 50:100] = 3.0e38                                                 # tiny norms (outside the fast range)
    # a norm with an all-ones significand (the one divisor class Markstein's theorem excludes): weights 1, 0.5 and
    # 0.5 - 2^-23 sum to 2 - 2^-23 = 0x3fffffff
    d2[:, 100:150] = torch.tensor([1.0, 2.0, 2.0 * (1 + 2.0 ** -22)])
    nbr = pkg.ops.NeighborIndex(idx.to(dev), M)
    out = pkg.ops.three_interpolate(feats.to(dev), nbr, d2.to(dev)).cpu()
    ref = canon.interp(feats, idx, d2)
    same = (out.view(torch.int32) == ref.view(torch.int32)) | (torch.isnan(out) & torch.isnan(ref))
    assert bool(same.all()), f"{int((~same).sum())} of {same.numel()} quotients differ"
    w = 1.0 / (d2 + 1e-9)
    norm_bits = ((w[..., 0] + w[..., 1]) + w[..., 2]).view(torch.int32) & 0x7fffff
    assert int((norm_bits == 0x7fffff).sum()) > 0, "no all-ones divisor in the fixture"


# --------------------------------------------------------------------------- K3/K4 knn (expanded form) + K9

@pytest.mark.parametrize("name", ["knn_F3", "knn_F64", "knn_F20"])
def test_knn_golden(pkg, dev, golden, name):
    g = golden(name)
    idx = pkg.dgcnn.knn(g["x"].to(dev), g["k"])
    assert idx.dtype == torch.int64 and torch.equal(idx.cpu(), g["idx"])


@pytest.mark.parametrize("F,N,k", [(3, 4096, 20), (64, 2048, 20), (64, 1000, 32), (128, 515, 16), (9, 300, 40),
                                   (256, 130, 8), (17, 100, 100), (6, 70, 5)])
def test_knn_vs_oracle_both_layouts(pkg, dev, F, N, k):
    x = torch.randn(2, F, N, generator=_gen(F + N))
    if F == 3:                                # S3DIS-like room offsets: heavy cancellation, many exact ties
        x = x * 0.3 + torch.tensor([17.0, 12.0, 1.5]).view(1, 3, 1)
    want = canon.knn_expand(x, k)[0]
    assert torch.equal(pkg.ops.knn_graph(x.to(dev), k).cpu(), want)
    xt = x.to(dev).transpose(1, 2).contiguous().transpose(1, 2)       # point-major memory, (B,F,N) view
    assert torch.equal(pkg.ops.knn_graph(xt, k).cpu(), want)


def test_knn_lattice_ties(pkg, dev):
    x = torch.randint(-127, 128, (2, 64, 512), generator=_gen(11)).float() / 256
    x[:, :, 300:] = x[:, :, :212]             # duplicated points: distance-0 ties
    assert torch.equal(pkg.ops.knn_graph(x.to(dev), 20).cpu(), canon.knn_expand(x, 20)[0])


def test_graph_feature_golden_and_backward(pkg, dev, golden):
    g = golden("graph_feature")
    xd = g["x"].to(dev).requires_grad_(True)
    out = pkg.dgcnn.get_graph_feature(xd, k=g["k"])
    assert out.shape == g["out"].shape and torch.equal(out.detach().cpu(), g["out"])
    w = torch.randn(out.shape, generator=_gen(6))
    (out * w.to(dev)).sum().backward()
    xr = g["x"].clone().requires_grad_(True)
    (O.get_graph_feature(xr, g["k"]) * w).sum().backward()
    assert torch.allclose(xd.grad.cpu(), xr.grad, rtol=RTOL, atol=ATOL)
    # explicit int64 idx, as EdgeConv callers may pass (dgcnn.py:24)
    idx = pkg.dgcnn.knn(g["x"].to(dev), g["k"])
    assert torch.equal(pkg.dgcnn.get_graph_feature(g["x"].to(dev), g["k"], idx=idx).cpu(), g["out"])


@pytest.mark.parametrize("F,N,k", [(3, 4096, 20), (64, 1024, 20), (100, 300, 7)])
def test_edge_feature_vs_oracle(pkg, dev, F, N, k):
    x = torch.randn(2, F, N, generator=_gen(F))
    idx = torch.randint(0, N, (2, N, k), generator=_gen(k))
    out = pkg.dgcnn.get_graph_feature(x.to(dev), k, idx=idx.to(dev))
    assert torch.equal(out.cpu(), canon.edge_feature(x, idx))


def test_error_behaviour(pkg, dev):
    x = torch.rand(1, 16, 3, device=dev)
    with pytest.raises(RuntimeError):                       # torch.topk raises when K > N (common.py:61)
        pkg.common.group(x[:, :2], x, x, 0.1, 17)
    with pytest.raises(RuntimeError):
        pkg.dgcnn.knn(torch.rand(1, 3, 10, device=dev), 11)
    with pytest.raises(TypeError):
        pkg.common.sample(x.double(), 4)


# --------------------------------------------------------------------------- K4 tensor-core kNN (tcgen05)

def _tc_debug(pkg, x, k, want_scores=False):
    from ctypes import c_void_p
    B, F, N = x.shape
    lib = pkg._lib
    idx = torch.empty(B, N, k, dtype=torch.int32, device=x.device)
    nb = lib.size("pcnbr_knn_expand_ws_bytes", B, F, N, k)
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
    scores = torch.full((B, N, N), float("nan"), device=x.device) if want_scores else None
    stats = torch.zeros(2, dtype=torch.int32, device=x.device)
    sb, sf, sn = x.stride()
    lib.call("pcnbr_knn_tc_debug_f32", x.data_ptr(), B, F, N, sf, sn, k, idx.data_ptr(), ws.data_ptr(), nb,
             scores.data_ptr() if want_scores else None, stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return idx, scores, stats.cpu()


def test_knn_tc_scores_are_lower_bounds_within_margin(pkg, dev):
    """The tcgen05 fp16 tile (TMA + bulk-copied tail slice + UMMA descriptors + TMEM readback) delivers, in pass 1,
    L_ij = a_i.a_j - |a_j|^2/2 - C1 |a_i||a_j| for the centred, power-of-two scaled features a = (x - mean)/S:
    a LOWER bound of the exact score that is at most 2 C1 |a_i||a_j| + C0 below it (C1 = 1e-3 covers the fp16
    rounding of both operands) -- with and without a large common offset."""
    for offset, scale in ((0.0, 1.0), (3.0, 0.1)):
        x = torch.randn(2, 64, 512, generator=_gen(21)) * scale + offset
        idx, scores, stats = _tc_debug(pkg, x.to(dev), 20, want_scores=True)
        xd = x.double()
        mean = xd.mean(dim=2, keepdim=True)
        v = (x.abs().amax(dim=(1, 2)) + mean.abs().amax(dim=(1, 2)).float()).double()
        S = 2.0 ** (torch.floor(torch.log2(v)) + 1)                   # next binade above max|x| + max|mean|
        a = (xd - mean) / S.view(-1, 1, 1)
        exact = torch.matmul(a.transpose(1, 2), a) - 0.5 * (a ** 2).sum(1).unsqueeze(1)
        norm = (a ** 2).sum(1).sqrt()
        pair = norm.unsqueeze(2) * norm.unsqueeze(1)
        bmax = norm.amax(dim=1).view(-1, 1, 1)
        c0 = 2e-5 * bmax ** 2 + 1e-6 * bmax + 1e-7
        gap = exact - scores.cpu().double()
        assert not torch.isnan(scores).any()
        assert (gap >= -c0).all(), f"not a lower bound: min gap {gap.min().item():.3e}"
        assert (gap <= 2.1e-3 * pair + c0).all(), f"max gap/pair {(gap / (pair + 1e-12)).max().item():.3e}"
        assert torch.equal(idx.cpu(), canon.knn_expand(x, 20)[0])
        assert stats[1] == 0 and stats[0] <= 2 * 512 * 48        # ~k + a few survivors per row, no overflow


@pytest.mark.parametrize("F,N,k,B", [(64, 4096, 20, 2), (64, 1000, 20, 3), (32, 2048, 16, 2), (64, 300, 32, 2), (64, 4096, 1, 1),
                                     (3, 4096, 20, 2), (6, 1024, 20, 2), (17, 600, 8, 2), (40, 512, 20, 2)])
def test_knn_tc_vs_oracle(pkg, dev, F, N, k, B):
    x = torch.randn(B, F, N, generator=_gen(F + N + k))
    if N == 1000:                                  # post-LeakyReLU-like features: common offset, small spread
        x = x * 0.05 + 1.0
    if F == 3:                                     # S3DIS xyz: 1 m x 1 m x 3 m block at a 17 m room offset
        x = torch.rand(B, F, N, generator=_gen(5)) * torch.tensor([1.0, 1.0, 3.0]).view(1, 3, 1) + torch.tensor([17.0, 12.0, 0.0]).view(1, 3, 1)
    want = canon.knn_expand(x, k)[0]
    idx, _, stats = _tc_debug(pkg, x.to(dev), k)
    assert torch.equal(idx.cpu(), want)
    xt = x.to(dev).transpose(1, 2).contiguous().transpose(1, 2)       # point-major input: no transposing copy
    assert torch.equal(pkg.ops.knn_graph(xt, k).cpu(), want)
    assert stats[1] == 0


def test_knn_tc_degenerate_rows_fall_back_to_exact_scan(pkg, dev):
    """Duplicated points / lattice ties blow the survivor queue for some rows: those rows are re-ranked by an
    exact full scan, the result must still be the canonical one."""
    x = torch.randint(-3, 4, (2, 64, 512), generator=_gen(33)).float() / 4
    x[:, :, 256:] = x[:, :, :256]
    want = canon.knn_expand(x, 20)[0]
    idx, _, stats = _tc_debug(pkg, x.to(dev), 20)
    assert torch.equal(idx.cpu(), want)
    x2 = torch.zeros(1, 64, 300)                   # all points identical: every row overflows
    idx2, _, stats2 = _tc_debug(pkg, x2.to(dev), 20)
    assert torch.equal(idx2.cpu(), canon.knn_expand(x2, 20)[0]) and stats2[1] == 300

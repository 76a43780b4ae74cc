"""GPU: parity at the sizes of BASELINE configs[3] (PointNeXt on 24 k-point chunks) and configs[4] (kernel sweep
N = 4 k - 100 k), covering every dispatch branch of the C ABI that the 4096-point tests do not reach:

  pcnbr_fps_f32      N <= 8192 register kernel | 4-CTA cluster (N <= 32768) | 8-CTA cluster (N <= 65536) | 16-CTA cluster
                     (N <= 131072) | global-memory kernel (larger N, or more clouds than the clusters can host)
  pcnbr_ball_query / pcnbr_knn_direct / pcnbr_group at N in {8192, 24000, 65536}
  pcnbr_knn_expand   tensor-core path for 4096 < N <= 65535 (16-bit survivor queue) and the CUDA-core path above it

Everything is compared bit-exactly with oracle/canon.c on the same seeded inputs
(/root/reference/models/utils/common.py:25-31, 54-61, 110-114; models/dgcnn/dgcnn.py:16-20)."""
import pytest
import torch

from oracle import canon, ref_ops as O

pytestmark = pytest.mark.gpu


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _chunk(B, N, seed):
    """S3DIS-shaped xyz for an N-point chunk: the same density as a 4096-point 1 m x 1 m block (a larger footprint)."""
    g = _gen(seed)
    side = (N / 4096.0) ** 0.5
    xy = torch.rand(B, N, 2, generator=g) * side + torch.randint(0, 20, (B, 1, 2), generator=g).float()
    z = torch.rand(B, N, 1, generator=g) * 3.0
    return torch.cat((xy, z), dim=2).contiguous()


@pytest.mark.parametrize("B,N,C,branch", [
    (8, 24000, 1024, "cluster4"),      # BASELINE configs[3]: PointNeXt sa1 on 24 k-point chunks
    (2, 32768, 128, "cluster4"),
    (2, 40000, 256, "cluster8"),
    (1, 65536, 64, "cluster8"),
    (2, 100000, 64, "cluster16"),      # configs[4] upper end: 16-CTA cluster (non-portable size)
    (1, 131072, 32, "cluster16"),
    (1, 140000, 24, "big"),            # beyond what a cluster's registers hold
    (40, 9000, 48, "big"),             # more clouds than the clusters can host at once
])
def test_fps_large_every_dispatch_branch(pkg, dev, B, N, C, branch):
    xyz = _chunk(B, N, seed=N + B)
    start = torch.randint(0, N, (B,), generator=_gen(C), dtype=torch.int32)
    pkg._lib.prof_enable(True)
    pkg._lib.prof_collect()
    idx, coords = pkg.ops.farthest_point_sample(xyz.to(dev), C, start.to(dev), return_coords=True)
    ran = pkg._lib.prof_collect()
    pkg._lib.prof_enable(False)
    want = {"cluster4": "fps_cluster_kernel", "cluster8": "fps_cluster_kernel", "cluster16": "fps_cluster_kernel", "big": "fps_big_kernel"}[branch]
    assert any(k.startswith(want) for k in ran), f"expected {want}, ran {sorted(ran)}"
    o_idx, o_coords = canon.fps(xyz, C, start)
    assert torch.equal(idx.cpu(), o_idx)
    assert torch.equal(coords.cpu(), o_coords)


@pytest.mark.parametrize("N", [8192, 24000, 65536])
@pytest.mark.parametrize("r,K", [(0.1, 32), (0.2, 16)])
def test_ball_query_and_group_large(pkg, dev, N, r, K):
    B, M, D = 2, 1024, 6
    xyz = _chunk(B, N, seed=N)
    start = torch.zeros(B, dtype=torch.int32)
    cen = canon.fps(xyz, M, start)[1]
    idx = pkg.ops.query_ball_point(r, K, xyz.to(dev), cen.to(dev))
    o_idx = canon.ball_query(cen, xyz, r, K)
    assert torch.equal(idx.cpu(), o_idx)
    feat = torch.randn(B, N, D, generator=_gen(N + K))
    fd = feat.to(dev).requires_grad_(True)
    out = pkg.ops.group_points(xyz.to(dev), fd, cen.to(dev), pkg.ops.NeighborIndex(idx, N), r)
    assert torch.equal(out.detach().cpu(), canon.group(cen, xyz, feat, o_idx, r, True))
    w = torch.randn(out.shape, generator=_gen(1))
    (out * w.to(dev)).sum().backward()
    # scatter-add backward against a float64 accumulation of the same terms.  The padded balls make the lowest indices hubs
    # (one gradient term from each of the 1024 centroids); where such a sum cancels, 1e-4 of the RESULT is below fp32
    # rounding of the terms (torch's own fp32 index_put is 1e-4 off there), so the bound is 1e-4 relative or 2e-7 of
    # sum |term| -- a few fp32 ulps per term.
    ref = torch.zeros(B, N, D, dtype=torch.float64)
    mag = torch.zeros(B, N, D, dtype=torch.float64)
    for b in range(B):
        terms = w[b].reshape(-1, 3 + D)[:, 3:].double()
        ref[b].index_add_(0, o_idx[b].reshape(-1).long(), terms)
        mag[b].index_add_(0, o_idx[b].reshape(-1).long(), terms.abs())
    err = (fd.grad.cpu().double() - ref).abs()
    assert bool((err <= 1e-4 * ref.abs() + 2e-7 * mag + 1e-12).all()), f"max err {err.max().item():.3e}"


@pytest.mark.parametrize("N", [8192, 24000, 65536])
def test_self_ball_query_large(pkg, dev, N):
    """InvResMLP's self-query (centroids == points, common.py:273) on a chunk-sized level."""
    xyz = _chunk(1, N, seed=7 * N)
    q = xyz[:, :2048].contiguous()
    assert torch.equal(pkg.ops.query_ball_point(0.1, 32, xyz.to(dev), q.to(dev)).cpu(), canon.ball_query(q, xyz, 0.1, 32))


@pytest.mark.parametrize("N,M", [(8192, 1024), (24000, 1024), (65536, 1024), (100000, 4096)])
def test_three_nn_and_interpolate_large(pkg, dev, N, M):
    B, D = 2, 32
    fine = _chunk(B, N, seed=N + 1)
    coarse = canon.fps(fine, M, torch.zeros(B, dtype=torch.int32))[1]
    idx, d2 = pkg.ops.knn_points(fine.to(dev), coarse.to(dev), 3)
    o_idx, o_d2 = canon.knn_direct(fine, coarse, 3)
    assert torch.equal(idx.cpu(), o_idx) and torch.equal(d2.cpu(), o_d2)
    feats = torch.randn(B, M, D, generator=_gen(M))
    out = pkg.common.interpolate(feats.to(dev), fine.to(dev), coarse.to(dev), 3)
    assert torch.equal(out.cpu(), canon.interp(feats, o_idx, o_d2))


@pytest.mark.parametrize("F,N,k", [(64, 8192, 20), (64, 16384, 20), (64, 6000, 32), (3, 24000, 16), (32, 12345, 20)])
def test_knn_graph_large_tensor_core_path(pkg, dev, F, N, k):
    """4096 < N <= 65535: the tcgen05 kernel with its 16-bit survivor queue (csrc/knn_tc.cu)."""
    x = torch.randn(1, F, N, generator=_gen(F + N))
    if F == 3:
        x = _chunk(1, N, seed=3).transpose(1, 2).contiguous()
    want = canon.knn_expand(x, k)[0]
    pkg._lib.prof_enable(True)
    pkg._lib.prof_collect()
    got = pkg.ops.knn_graph(x.to(dev), k)
    ran = pkg._lib.prof_collect()
    pkg._lib.prof_enable(False)
    assert any(name.startswith("knn_tc_kernel") for name in ran), sorted(ran)
    assert torch.equal(got.cpu(), want)


def test_knn_graph_baseline_shape_golden_from_reference(pkg, dev, golden):
    """BASELINE configs[1] shape (F=64, N=4096, k=20): the indices of the UNMODIFIED reference's knn()
    (models/dgcnn/dgcnn.py:7-21, generated by oracle/make_golden_large.py) -- the headline configuration is pinned by the
    reference itself, not only by the port."""
    g = golden("knn_F64_N4096")
    x = torch.randn(1, 64, 4096, generator=_gen(int(g["seed"])))
    idx = pkg.dgcnn.knn(x.to(dev), int(g["k"]))
    assert torch.equal(idx.cpu().to(torch.int16), g["idx"].to(torch.int16))


def test_knn_graph_cuda_core_path_above_65535(pkg, dev):
    """N > 65535 (the 16-bit survivor queue does not apply): the exact CUDA-core selection.  The oracle evaluates a
    spread of query rows (the full (N,N) scan is 4.9e9 pairs)."""
    F, N, k = 8, 70000, 16
    x = torch.randn(1, F, N, generator=_gen(99))
    got = pkg.ops.knn_graph(x.to(dev), k).cpu()
    rows = torch.cat((torch.arange(0, 64), torch.arange(N - 64, N), torch.randint(0, N, (384,), generator=_gen(1)))).int()
    want = canon.knn_expand_rows(x[0], k, rows)
    assert torch.equal(got[0, rows.long()], want)


def test_large_goldens_from_reference(pkg, dev, golden):
    """Outputs of the UNMODIFIED reference at sizes above 4096 points (oracle/make_golden_large.py): sample() on a
    24 000-point chunk, group() and interpolate() over 8192 source points."""
    g = golden("fps_24k")
    xyz = _chunk(1, g["N"], g["seed"])
    assert torch.equal(pkg.common.sample(xyz.to(dev), g["C"], g["start"].to(dev)).cpu(), g["coords"])
    g = golden("group_8k")
    gen = _gen(g["seed"])
    p = torch.rand(1, 8192, 3, generator=gen) * 0.5
    feat = torch.randn(1, 8192, 6, generator=gen)
    out = pkg.common.group(g["centroids"].to(dev), p.to(dev), feat.to(dev), g["r"], g["K"], True)
    assert torch.equal(out.cpu(), g["out"])
    h = golden("interp_8k")
    up = pkg.common.interpolate(h["points"].to(dev), p.to(dev), g["centroids"].to(dev)).cpu()
    assert torch.equal(up[:, :512], h["out_first"]) and torch.equal(up.double().sum(dim=1), h["out_sum"])


# --------------------------------------------------------------------------- cell-grid selection == M x N scan (csrc/grid.cu)

def _brute_ball(pkg, q, p, r, K):
    B, M, N = q.shape[0], q.shape[1], p.shape[1]
    idx = torch.empty(B, M, K, dtype=torch.int32, device=q.device)
    pkg._lib.call("pcnbr_ball_query_f32", q.data_ptr(), p.data_ptr(), B, M, N, pkg.ops._r2(r), K, idx.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
    return idx


def _grid_ball(pkg, q, p, r, K):
    B, M, N = q.shape[0], q.shape[1], p.shape[1]
    idx = torch.empty(B, M, K, dtype=torch.int32, device=q.device)
    nb = pkg._lib.size("pcnbr_grid_ws_bytes", B, N)
    ws = torch.empty(nb, dtype=torch.uint8, device=q.device)
    pkg._lib.call("pcnbr_ball_query_grid_f32", q.data_ptr(), p.data_ptr(), B, M, N, pkg.ops._r2(r), K, idx.data_ptr(), ws.data_ptr(), nb,
                  torch.cuda.current_stream().cuda_stream)
    return idx


def _knn_both(pkg, q, p, k):
    B, M, N = q.shape[0], q.shape[1], p.shape[1]
    out = []
    for name in ("pcnbr_knn_direct_f32", "pcnbr_knn_direct_grid_f32"):
        idx = torch.empty(B, M, k, dtype=torch.int32, device=q.device)
        d2 = torch.empty(B, M, k, dtype=torch.float32, device=q.device)
        if name.endswith("grid_f32"):
            nb = pkg._lib.size("pcnbr_grid_ws_bytes", B, N)
            ws = torch.empty(nb, dtype=torch.uint8, device=q.device)
            pkg._lib.call(name, q.data_ptr(), p.data_ptr(), B, M, N, k, idx.data_ptr(), d2.data_ptr(), ws.data_ptr(), nb,
                          torch.cuda.current_stream().cuda_stream)
        else:
            pkg._lib.call(name, q.data_ptr(), p.data_ptr(), B, M, N, k, idx.data_ptr(), d2.data_ptr(), torch.cuda.current_stream().cuda_stream)
        out.append((idx, d2))
    return out


def _clouds(kind, B, N, seed):
    g = _gen(seed)
    if kind == "block":                       # S3DIS-shaped block at a room offset
        return _chunk(B, N, seed)
    if kind == "flat":                        # a wall: zero extent along y
        p = _chunk(B, N, seed)
        p[:, :, 1] = 7.25
        return p
    if kind == "lattice":                     # exact ties in d2, duplicated points
        p = torch.randint(-40, 41, (B, N, 3), generator=g).float() / 64
        p[:, N // 2:] = p[:, : N - N // 2]
        return p
    if kind == "clustered":                   # 90 % of the points in 1 % of the volume
        p = torch.rand(B, N, 3, generator=g)
        p[:, : int(0.9 * N)] = p[:, : int(0.9 * N)] * 0.05 + 0.4
        return p
    if kind == "point":                       # every point identical
        return torch.full((B, N, 3), 1.5)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["block", "flat", "lattice", "clustered", "point"])
@pytest.mark.parametrize("N,M,r,K", [(4096, 1024, 0.1, 32), (2500, 300, 0.2, 16), (5000, 64, 0.05, 64), (3000, 100, 0.3, 100)])
def test_grid_ball_query_equals_brute_force(pkg, dev, kind, N, M, r, K):
    p = _clouds(kind, 2, N, seed=N + M).to(dev)
    q = p[:, torch.randperm(N, generator=_gen(1))[:M]].contiguous()
    q[:, : M // 8] += 0.37                      # some queries off the points, a few outside the bounding box
    q[:, 0] = torch.tensor([-50.0, 80.0, 3.0], device=dev)
    assert torch.equal(_grid_ball(pkg, q, p, r, K), _brute_ball(pkg, q, p, r, K))


@pytest.mark.parametrize("kind", ["block", "flat", "lattice", "clustered", "point"])
@pytest.mark.parametrize("N,M,k", [(1024, 4096, 3), (600, 2000, 8), (4096, 1000, 20), (700, 300, 32), (50, 64, 3)])
def test_grid_knn_equals_brute_force(pkg, dev, kind, N, M, k):
    p = _clouds(kind, 2, N, seed=N + k).to(dev)
    q = _clouds(kind, 2, M, seed=M + k + 1).to(dev)
    q[:, 0] = torch.tensor([-50.0, 80.0, 3.0], device=dev)
    (bi, bd), (gi, gd) = _knn_both(pkg, q, p, k)
    assert torch.equal(gi, bi) and torch.equal(gd, bd)

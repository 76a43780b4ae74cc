"""GPU: out-of-bounds WRITE guards for the C ABI (compute-sanitizer is closed on this GPU pool, profiles/r4_sanitizer_closed.txt).

Every output buffer handed to an entry point is a window inside a larger arena filled with a sentinel byte pattern; after
the call the red zones before and after the window must be untouched, and the window itself must not contain the sentinel
where a result is expected.  Sizes are ragged on purpose (tails of every tile / vector width).  The values are checked
against the oracle elsewhere; this file only proves that no kernel writes outside the extents the header declares."""
import pytest
import torch

pytestmark = pytest.mark.gpu

PAD = 4096          # bytes of red zone on either side
SENT = 0xA5


class Arena:
    def __init__(self, dev):
        self.dev, self.items = dev, []

    def out(self, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        nbytes_al = (nbytes + 15) // 16 * 16
        raw = torch.full((PAD + nbytes_al + PAD,), SENT, dtype=torch.uint8, device=self.dev)
        win = raw[PAD:PAD + nbytes].view(dtype).view(*shape)
        self.items.append((raw, nbytes))
        return win

    def check(self):
        torch.cuda.synchronize()
        for i, (raw, nbytes) in enumerate(self.items):
            assert bool((raw[:PAD] == SENT).all()), f"buffer {i}: write BEFORE the buffer"
            assert bool((raw[PAD + nbytes:] == SENT).all()), f"buffer {i}: write PAST the end of the buffer"


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cloud(B, N, seed, dev):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, N, 3, generator=g) * torch.tensor([1.0, 1.0, 3.0]) + torch.tensor([7.0, 3.0, 0.0])).to(dev)


@pytest.mark.parametrize("B,N,C", [(3, 1000, 77), (2, 4097, 130), (1, 9001, 33), (2, 33000, 17)])
def test_fps_writes_stay_in_bounds(pkg, dev, B, N, C):
    L, a = pkg._lib, Arena(dev)
    xyz = _cloud(B, N, N, dev)
    start = torch.zeros(B, dtype=torch.int32, device=dev)
    idx, out = a.out((B, C), torch.int32), a.out((B, C, 3), torch.float32)
    nb = L.size("pcnbr_fps_ws_bytes", B, N)
    ws = a.out((max(nb, 4),), torch.uint8)
    L.call("pcnbr_fps_f32", xyz.data_ptr(), B, N, C, start.data_ptr(), idx.data_ptr(), out.data_ptr(), ws.data_ptr(), nb, _stream())
    a.check()
    assert int(idx.min()) >= 0 and int(idx.max()) < N


@pytest.mark.parametrize("N,M,K", [(2049, 301, 32), (5000, 77, 17), (777, 33, 40), (3001, 130, 65)])
def test_selection_writes_stay_in_bounds(pkg, dev, N, M, K):
    L, a = pkg._lib, Arena(dev)
    B = 2
    p = _cloud(B, N, N + M, dev)
    q = p[:, :M].contiguous()
    r2 = pkg.ops._r2(0.15)
    i1 = a.out((B, M, K), torch.int32)
    L.call("pcnbr_ball_query_f32", q.data_ptr(), p.data_ptr(), B, M, N, r2, K, i1.data_ptr(), _stream())
    i2 = a.out((B, M, K), torch.int32)
    nb = L.size("pcnbr_grid_ws_bytes", B, N)
    ws = a.out((nb,), torch.uint8)
    L.call("pcnbr_ball_query_grid_f32", q.data_ptr(), p.data_ptr(), B, M, N, r2, K, i2.data_ptr(), ws.data_ptr(), nb, _stream())
    k = min(K, 32)
    i3, d3 = a.out((B, M, k), torch.int32), a.out((B, M, k), torch.float32)
    L.call("pcnbr_knn_direct_f32", q.data_ptr(), p.data_ptr(), B, M, N, k, i3.data_ptr(), d3.data_ptr(), _stream())
    i4, d4 = a.out((B, M, k), torch.int32), a.out((B, M, k), torch.float32)
    ws2 = a.out((nb,), torch.uint8)
    L.call("pcnbr_knn_direct_grid_f32", q.data_ptr(), p.data_ptr(), B, M, N, k, i4.data_ptr(), d4.data_ptr(), ws2.data_ptr(), nb, _stream())
    a.check()
    assert torch.equal(i1, i2) and torch.equal(i3, i4) and torch.equal(d3, d4)
    for t in (i1, i3):
        assert int(t.min()) >= 0 and int(t.max()) < N


@pytest.mark.parametrize("F,N,k", [(64, 1000, 20), (3, 4097, 20), (17, 515, 8), (100, 300, 7)])
def test_knn_graph_and_edge_feature_writes_stay_in_bounds(pkg, dev, F, N, k):
    L, a = pkg._lib, Arena(dev)
    B = 2
    x = torch.randn(B, F, N, generator=torch.Generator().manual_seed(F + N)).to(dev)
    idx = a.out((B, N, k), torch.int32)
    nb = L.size("pcnbr_knn_expand_ws_bytes", B, F, N, k)
    ws = a.out((nb,), torch.uint8)
    L.call("pcnbr_knn_expand_f32", x.data_ptr(), B, F, N, N, 1, k, idx.data_ptr(), ws.data_ptr(), nb, _stream())
    xt = x.transpose(1, 2).contiguous()
    ef = a.out((B, N, k, 2 * F), torch.float32)
    L.call("pcnbr_edge_feature_f32", xt.data_ptr(), idx.data_ptr(), B, N, F, k, ef.data_ptr(), _stream())
    off, perm = a.out((B, N + 1), torch.int32), a.out((B, N * k), torch.int32)
    nb2 = L.size("pcnbr_csr_rows_ws_bytes", B, N, k, N)
    ws2 = a.out((nb2,), torch.uint8)
    L.call("pcnbr_csr_build_rows", idx.data_ptr(), B, N, k, N, off.data_ptr(), perm.data_ptr(), ws2.data_ptr(), nb2, _stream())
    g = torch.randn(B, N, k, 2 * F, device=dev)
    gx = a.out((B, N, F), torch.float32)
    L.call("pcnbr_edge_feature_bwd_f32", g.data_ptr(), off.data_ptr(), perm.data_ptr(), B, N, F, k, gx.data_ptr(), _stream())
    a.check()
    assert int(idx.min()) >= 0 and int(idx.max()) < N and int(perm.max()) < N * k


@pytest.mark.parametrize("N,M,K,D", [(1000, 77, 32, 6), (513, 33, 16, 64), (300, 20, 7, 131), (2049, 100, 32, 0)])
def test_group_pool_interp_writes_stay_in_bounds(pkg, dev, N, M, K, D):
    L, a = pkg._lib, Arena(dev)
    B = 2
    p = _cloud(B, N, N + D, dev)
    q = p[:, :M].contiguous()
    feat = torch.randn(B, N, max(D, 1), device=dev)[:, :, :D].contiguous()
    idx = pkg.ops.query_ball_point(0.2, K, p, q)
    W = 3 + D
    for pitch in (W, (W + 3) // 4 * 4 if W > 32 else 32):
        out = a.out((B, M, K, pitch), torch.float32)
        L.call("pcnbr_group_f32", p.data_ptr(), feat.data_ptr() if D else None, q.data_ptr(), idx.data_ptr(), B, N, M, K, D, 0.2,
               out.data_ptr(), pitch, _stream())
    if D:
        nbr = pkg.ops.NeighborIndex(idx, N)
        off, perm = nbr.csr()
        gout = torch.randn(B, M, K, W, device=dev)
        gfeat = a.out((B, N, D), torch.float32)
        L.call("pcnbr_group_bwd_f32", gout.data_ptr(), W, off.data_ptr(), perm.data_ptr(), B, N, M * K, D, gfeat.data_ptr(), _stream())
        x = torch.randn(B, M, K, D, device=dev)
        if K <= 255:
            po, pa = a.out((B * M * D,), torch.float32), a.out((B * M * D,), torch.uint8)
            L.call("pcnbr_maxpool_f32", x.data_ptr(), B * M, K, D, K * D, D, 1, po.data_ptr(), pa.data_ptr(), _stream())
        i3, d3 = pkg.ops.knn_points(p, q, 3)
        coarse = torch.randn(B, M, D, device=dev)
        up, coef = a.out((B, N, D), torch.float32), a.out((B, N, 3), torch.float32)
        L.call("pcnbr_interp_f32", coarse.data_ptr(), i3.data_ptr(), d3.data_ptr(), B, N, M, D, 3, up.data_ptr(), coef.data_ptr(), _stream())
    a.check()


@pytest.mark.parametrize("M,N,K", [(300, 72, 100), (130, 260, 4100), (1000, 36, 2048), (4097, 516, 388)])
@pytest.mark.parametrize("kernel", ["3xtf32", "fp16split"])
def test_gemm_writes_stay_in_bounds(pkg, dev, M, N, K, kernel):
    """TMA-store epilogue with ragged M / N (clipped by the tensor map), split-K partials + reduction, both GEMM kernels."""
    a = Arena(dev)
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(dev)
    Bm = torch.randn(N, K, generator=g).to(dev)
    out = a.out((M, N), torch.float32)
    pkg.ops._gemm3x(A, False, Bm, False, M, N, K, out=out, force_h2=(kernel == "fp16split"))
    a.check()
    ref = A.double() @ Bm.double().t()
    assert float((out.double() - ref).abs().max()) <= 3e-5 * float(ref.abs().max())


@pytest.mark.parametrize("N,M,K", [(2049, 301, 32), (777, 33, 17), (9001, 130, 20)])
def test_length_aware_entry_points_write_in_bounds_and_in_range(pkg, dev, N, M, K):
    """pcnbr_*_len_f32: ragged per-cloud lengths, outputs in canary arenas; every index stays below the cloud's
    own length for the real rows and is the in-range filler for the padding rows."""
    L, a = pkg._lib, Arena(dev)
    B = 3
    p = _cloud(B, N, N + K, dev)
    lens = torch.tensor([N, max(K, N // 3), max(K, 41)], dtype=torch.int32, device=dev)
    for b in range(B):
        p[b, int(lens[b]):] = 0.0
    start = torch.tensor([N - 1, 5, 10 ** 6], dtype=torch.int32, device=dev)              # out-of-range start: clamped
    C = 64
    fi, fo = a.out((B, C), torch.int32), a.out((B, C, 3), torch.float32)
    nb = L.size("pcnbr_fps_ws_bytes", B, N)
    ws = a.out((max(nb, 4),), torch.uint8)
    L.call("pcnbr_fps_len_f32", p.data_ptr(), B, N, C, start.data_ptr(), lens.data_ptr(), fi.data_ptr(), fo.data_ptr(), ws.data_ptr(), nb, _stream())
    q = fo[:, :min(M, C)].contiguous()
    Mq = q.shape[1]
    qlen = torch.tensor([Mq, Mq // 2, 1], dtype=torch.int32, device=dev)
    gnb = L.size("pcnbr_grid_ws_bytes", B, N)
    for use_grid in (False, True):
        wsg = a.out((gnb,), torch.uint8) if use_grid else None
        bi = a.out((B, Mq, K), torch.int32)
        L.call("pcnbr_ball_query_len_f32", q.data_ptr(), p.data_ptr(), B, Mq, N, pkg.ops._r2(0.2), K, qlen.data_ptr(), lens.data_ptr(),
               bi.data_ptr(), wsg.data_ptr() if use_grid else None, gnb if use_grid else 0, _stream())
        ki, kd = a.out((B, Mq, 3), torch.int32), a.out((B, Mq, 3), torch.float32)
        L.call("pcnbr_knn_direct_len_f32", q.data_ptr(), p.data_ptr(), B, Mq, N, 3, qlen.data_ptr(), lens.data_ptr(), ki.data_ptr(),
               kd.data_ptr(), wsg.data_ptr() if use_grid else None, gnb if use_grid else 0, _stream())
        a.check()
        for b in range(B):
            assert int(bi[b].max()) < int(lens[b]) and int(ki[b].max()) < int(lens[b]) and int(bi[b].min()) >= 0
    assert all(int(fi[b].max()) < int(lens[b]) for b in range(B))
    F = 64 if N < 5000 else 3
    x = torch.randn(B, F, N, generator=torch.Generator().manual_seed(N)).to(dev)
    gi = a.out((B, N, K), torch.int32)
    nb2 = L.size("pcnbr_knn_expand_ws_bytes", B, F, N, K)
    ws2 = a.out((nb2,), torch.uint8)
    L.call("pcnbr_knn_expand_len_f32", x.data_ptr(), B, F, N, N, 1, K, lens.data_ptr(), gi.data_ptr(), ws2.data_ptr(), nb2, _stream())
    a.check()
    for b in range(B):
        n = int(lens[b])
        assert int(gi[b, :n].max()) < n and int(gi[b].min()) >= 0 and int(gi[b, n:].max() if n < N else 0) < K


@pytest.mark.parametrize("N,E,D", [(1000, 777, 6), (333, 4100, 130), (64, 33, 3)])
def test_gather_rows_and_square_distance_write_in_bounds(pkg, dev, N, E, D):
    L, a = pkg._lib, Arena(dev)
    B = 2
    src = torch.randn(B, N, D, device=dev)
    idx = torch.randint(-5, N + 5, (B, E), device=dev, dtype=torch.int32)                  # out-of-range indices are clamped
    out = a.out((B, E, D), torch.float32)
    L.call("pcnbr_gather_rows_f32", src.data_ptr(), idx.data_ptr(), B, N, E, D, out.data_ptr(), _stream())
    a.check()
    assert torch.equal(out, src[torch.arange(B, device=dev)[:, None], idx.clamp(0, N - 1).long()])
    nbr = pkg.ops.NeighborIndex(idx.clamp(0, N - 1), N)
    off, perm = nbr.csr()
    gs = a.out((B, N, D), torch.float32)
    L.call("pcnbr_gather_rows_bwd_f32", out.data_ptr(), off.data_ptr(), perm.data_ptr(), B, N, E, D, gs.data_ptr(), _stream())
    p, q = _cloud(B, N, N, dev), _cloud(B, E, E, dev)
    d2 = a.out((B, N, E), torch.float32)
    L.call("pcnbr_square_distance_f32", p.data_ptr(), q.data_ptr(), B, N, E, d2.data_ptr(), _stream())
    a.check()
    assert bool(torch.isfinite(gs).all()) and bool((d2 >= 0).all())

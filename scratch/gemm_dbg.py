"""Debug driver for the 3xTF32 tcgen05 GEMM: accuracy vs fp64 and timing vs the library fp32 SGEMM."""
import sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
pkg = ge.load_package(); lib = pkg._lib; lib.load()
dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False
st = lambda: torch.cuda.current_stream().cuda_stream

def split(x, want=True, wantT=False):
    R, C = x.shape
    hi = torch.empty_like(x) if want else None; lo = torch.empty_like(x) if want else None
    hiT = torch.empty(C, R, device=x.device) if wantT else None; loT = torch.empty(C, R, device=x.device) if wantT else None
    p = lambda t: t.data_ptr() if t is not None else None
    lib.call("pcnbr_split_tf32", x.data_ptr(), R, C, p(hi), p(lo), p(hiT), p(loT), st())
    return hi, lo, hiT, loT

def gemm(ah, al, bh, bl, bias=None, splits=None):
    M, K = ah.shape; N = bh.shape[0]
    if splits is None: splits = lib.size("pcnbr_gemm3x_splits", M, N, K)
    nb = lib.size("pcnbr_gemm3x_ws_bytes", M, N, K, splits)
    ws = torch.empty(max(nb, 4), dtype=torch.uint8, device=ah.device)
    C = torch.empty(M, N, device=ah.device)
    lib.call("pcnbr_gemm3x_f32", ah.data_ptr(), al.data_ptr(), bh.data_ptr(), bl.data_ptr(), M, N, K,
             bias.data_ptr() if bias is not None else None, C.data_ptr(), splits, ws.data_ptr(), nb, st())
    return C, splits

g = torch.Generator().manual_seed(0)
for (M, N, K, use_bias) in [(1000, 200, 100, True), (65536, 512, 1408, True), (65536, 1024, 384, False), (65536, 256, 512, True),
                            (65536, 128, 64, False), (512, 1408, 65536, False), (1024, 384, 65536, False), (128, 64, 65536, False), (300, 130, 36, True)]:
    A = (torch.randn(M, K, generator=g) * 1.3 + 0.2).to(dev); Bm = torch.randn(N, K, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev) if use_bias else None
    ah, al, _, _ = split(A); bh, bl, _, _ = split(Bm)
    # split correctness: hi + lo ~= x to 2^-21
    assert ((ah + al - A).abs() <= A.abs() * 2.0 ** -21 + 1e-30).all(), "split residual"
    splits = None
    if bias is not None and lib.size("pcnbr_gemm3x_splits", M, N, K) > 1: splits = 1
    C, sp = gemm(ah, al, bh, bl, bias, splits)
    torch.cuda.synchronize()
    sub = slice(0, min(M, 4096))
    ref = A[sub].double() @ Bm.double().t() + (bias.double() if bias is not None else 0)
    scale = (A[sub].double().abs() @ Bm.double().abs().t()) + 1e-30
    err = ((C[sub].double() - ref).abs() / scale).max().item()
    lib32 = A[sub] @ Bm.t() + (bias if bias is not None else 0)
    err32 = ((lib32.double() - ref).abs() / scale).max().item()
    # also check the tail rows / transposes path
    tail = slice(max(0, M - 300), M)
    ref_t = A[tail].double() @ Bm.double().t() + (bias.double() if bias is not None else 0)
    err_t = ((C[tail].double() - ref_t).abs() / ((A[tail].double().abs() @ Bm.double().abs().t()) + 1e-30)).max().item()
    # timing
    for _ in range(2): gemm(ah, al, bh, bl, bias, splits)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5): gemm(ah, al, bh, bl, bias, splits)
    ev1.record(); torch.cuda.synchronize(); t_ours = ev0.elapsed_time(ev1) / 5
    for _ in range(2): torch.addmm(bias, A, Bm.t()) if bias is not None else A @ Bm.t()
    ev0.record()
    for _ in range(5): torch.addmm(bias, A, Bm.t()) if bias is not None else A @ Bm.t()
    ev1.record(); torch.cuda.synchronize(); t_lib = ev0.elapsed_time(ev1) / 5
    ev0.record()
    for _ in range(5): split(A, True, True)
    ev1.record(); torch.cuda.synchronize(); t_split = ev0.elapsed_time(ev1) / 5
    print(f"M={M} N={N} K={K} splits={sp}: err/|a||b| ours {err:.2e} (tail {err_t:.2e}) lib-fp32 {err32:.2e} | ours {t_ours:.3f} ms = {2e-9 * M * N * K / t_ours:.1f} TFLOP/s, "
          f"library {t_lib:.3f} ms = {2e-9 * M * N * K / t_lib:.1f} TFLOP/s, split(A,+T) {t_split:.3f} ms")
# transposes
X = torch.randn(777, 130, generator=g).to(dev)
h, l, hT, lT = split(X, True, True)
assert torch.equal(hT, h.t().contiguous()) and torch.equal(lT, l.t().contiguous()), "transpose outputs"
print("split transposes ok")

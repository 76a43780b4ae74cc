"""All four operand layouts of pcnbr_gemm3x_f32 against float64; prints error and a coarse error map."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
dev = torch.device("cuda:0")
pad = lambda t: torch.nn.functional.pad(t, (0, (-t.shape[1]) % 4))
for (M, N, K) in [(128, 128, 32), (128, 32, 8), (256, 256, 64), (300, 72, 100), (64, 12, 4096)]:
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().t()
    for a_mn in (False, True):
        for b_mn in (False, True):
            Am = pad(A.t().contiguous() if a_mn else A).to(dev)
            Bm = pad(B.t().contiguous() if b_mn else B).to(dev)
            out = pkg.ops._gemm3x(Am, a_mn, Bm, b_mn, M, N, K).cpu().double()
            err = (out - ref).abs()
            print(f"M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)}: max err {err.max().item():.3e} (scale {ref.abs().max().item():.2f})"
                  f" bad rows {int((err.max(1).values > 1e-3).sum())}/{M} bad cols {int((err.max(0).values > 1e-3).sum())}/{N}")

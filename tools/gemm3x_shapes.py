"""The 3xTF32 GEMMs of PointNet++ SSG (32 x 4096) -- forward, input gradient and weight gradient of representative layers as
the layer functions issue them --, timed one by one (L2 flushed) against their HBM / tensor bounds, with the wait-time trace of
the kernel's warp roles (pcnbr_gemm2h_trace fills the same slots for gemm3x_kernel; slot 1 = the MMA warp waiting for the TMA).
    python tools/gemm3x_shapes.py [--trace]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package(); ops = pkg.ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
trace = "--trace" in sys.argv
tbuf = torch.zeros(148, 16, dtype=torch.int64, device=dev)
SLOTS = ["tma:ring", "mma:tma", "mma:acc", "mma:conv", "epi:full", "epi:slab", "cta", "conv:busy"]

def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 1):
        for _ in range(6):                  # L2 flush, and ~0.3 ms of GPU work in front of e0: the host-side part of the call
            flush.zero_()                   # (tensor-map encoding, ctypes) runs while the GPU is still busy, not inside [e0, e1]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts[1:])[len(ts[1:]) // 2]

def traced(fn):
    tbuf.zero_(); flush.zero_(); torch.cuda.synchronize()
    pkg._lib.call("pcnbr_gemm2h_trace", tbuf.data_ptr())
    fn(); torch.cuda.synchronize()
    pkg._lib.call("pcnbr_gemm2h_trace", None)
    t = tbuf.cpu().double(); t = t[t[:, 6] > 0]
    if len(t) == 0: return "-"
    life = t[:, 6].mean().item()
    return (f"{len(t)} CTAs {t[:, 8].mean().item() / 1e3:.0f} us (max {t[:, 8].max().item() / 1e3:.0f}): " +
            " ".join(f"{n} {t[:, i].mean().item() / life:.2f}" for i, n in enumerate(SLOTS) if i != 6))

print("| layer | GEMM | us | HBM us | tensor us (3 x TF32) | frac of the binding bound |")
print("|---|---|---:|---:|---:|---:|")
for name, R, Cin, Cout in [("SA1 l2", 1 << 20, 32, 32), ("SA1 l3", 1 << 20, 32, 64), ("SA2 l2", 1 << 18, 64, 64), ("SA2 l3", 1 << 18, 64, 128),
                           ("SA3 l3", 1 << 16, 128, 256), ("FP1", 1 << 17, 128, 128), ("FP2 l1", 1 << 15, 320, 256), ("SA4 l3", 1 << 14, 256, 512)]:
    x = torch.randn(R, Cin, generator=g).to(dev); w = (torch.randn(Cout, Cin, generator=g) / Cin ** 0.5).to(dev); gy = torch.randn(R, Cout, generator=g).to(dev)
    for what, fn in (("y = x W^T", lambda: ops._gemm3x(x, False, w, False, R, Cout, Cin)),
                     ("dx = gy W", lambda: ops._gemm3x(gy, False, w, True, R, Cin, Cout)),
                     ("dW = gy^T x", lambda: ops._wgrad3x(gy, x))):
        us = timed(fn)
        hbm = 4.0 * (R * Cin + R * Cout + Cin * Cout) / 6551e9 * 1e6
        ten = 3 * 2.0 * R * Cin * Cout / 678.35e12 * 1e6
        print(f"| {name} {R}x{Cin}->{Cout} | {what} | {us:.1f} | {hbm:.1f} | {ten:.1f} | {max(hbm, ten) / us:.2f} |" + (f" {traced(fn)}" if trace else ""))

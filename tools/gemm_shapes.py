"""The GEMMs of DGCNNWithColor's head (conv5 384 -> 1024, conv6 [384 | 1024] -> 512, conv7 512 -> 256 over 65536 rows) exactly
as the layer functions issue them on the fp16-split path: forward, input gradient(s) and weight gradient(s), timed one by one
with CUDA events (L2 flushed between launches), against the f16-pipe time of the three products each one issues.
    python tools/gemm_shapes.py [--reps N]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package(); ops = pkg.ops
dev = torch.device("cuda:0")
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 5
g = torch.Generator().manual_seed(0)
R = 65536
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
def rnd(*shape): return torch.randn(*shape, generator=g).to(dev)

def timed(fn):
    ts = []
    for _ in range(reps + 1):
        for _ in range(6):                  # L2 flush, and ~0.3 ms of GPU work in front of e0: the host-side part of the call
            flush.zero_()                   # (tensor-map encoding, ctypes) runs while the GPU is still busy, not inside [e0, e1]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts[1:])[len(ts[1:]) // 2]

trace = "--trace" in sys.argv
tbuf = torch.zeros(148, 16, dtype=torch.int64, device=dev)
SLOTS = ["tma:ring", "conv:tma", "mma:acc", "mma:conv", "epi:full", "epi:slab", "cta", "conv:busy"]
def traced(fn):
    """mean over the CTAs of each role's waiting time as a fraction of the CTA lifetime"""
    tbuf.zero_(); flush.zero_(); torch.cuda.synchronize()
    pkg._lib.call("pcnbr_gemm2h_trace", tbuf.data_ptr())
    fn(); torch.cuda.synchronize()
    pkg._lib.call("pcnbr_gemm2h_trace", None)
    t = tbuf.cpu().double(); t = t[t[:, 6] > 0]
    life = t[:, 6].mean().item()
    mhz = (t[:, 6] / t[:, 8] * 1e3).mean().item()
    return (f"{t[:, 8].mean().item() / 1e3:.0f} us (max {t[:, 8].max().item() / 1e3:.0f}) at {mhz:.0f} MHz: " +
            " ".join(f"{n} {t[:, i].mean().item() / life:.2f}" for i, n in enumerate(SLOTS) if i != 6))

rows = []
def layer(name, Cin, Cout):
    x, w, gy = rnd(R, Cin), rnd(Cout, Cin) / Cin ** 0.5, rnd(R, Cout)
    ax, aw, ag = ops._absmax(x), ops._absmax(w), ops._absmax(gy)
    wf, wt = ops._wsplit(w, False, aw), ops._wsplit(w, True, aw)
    gf = 2.0 * R * Cin * Cout
    planes = "--no-planes" not in sys.argv      # the layer path: forward / input gradient also write the operand planes the weight gradient reads
    xp, gp = (ops._new_planes(x, ax), ops._new_planes(gy, ag)) if planes else (None, None)
    for what, fn in (("y = x W^T", lambda: ops._gemm3x(x, False, w, False, R, Cout, Cin, amax_a=ax, amax_b=aw, b_split=wf, a_planes_out=xp)),
                     ("dx = gy W", lambda: ops._gemm3x(gy, False, w, True, R, Cin, Cout, amax_a=ag, amax_b=aw, b_split=wt, a_planes_out=gp)),
                     ("dW = gy^T x", lambda: ops._gemm3x(gy, True, x, True, Cout, Cin, R, amax_a=ag, amax_b=ax, a_mns=gp, b_mns=xp))):
        us = timed(fn)
        if trace: print(f"{name:13s} {what:12s} {us:6.1f} us | CTA {traced(fn)}")
        ideal = 3.0 * gf / 1356.7e12 * 1e6
        rows.append((name, what, Cin, Cout, us, ideal))

layer("conv5", 384, 1024)
layer("conv6[r_cat]", 384, 512)
layer("conv6[r5]", 1024, 512)
layer("conv7", 512, 256)
print("| layer | GEMM | Cin | Cout | us | f16-pipe time of the 3 products (us) | issued frac |")
print("|---|---|---:|---:|---:|---:|---:|")
for name, what, Cin, Cout, us, ideal in rows:
    print(f"| {name} | {what} | {Cin} | {Cout} | {us:.1f} | {ideal:.1f} | {ideal / us:.2f} |")
print(f"total {sum(r[4] for r in rows):.0f} us, pipe time {sum(r[5] for r in rows):.0f} us")

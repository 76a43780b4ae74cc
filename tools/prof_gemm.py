"""The three GEMMs of DGCNN's conv6 (65536 x 1408 -> 512: forward, input gradient, weight gradient) as the layer functions
issue them on the fp16-split path -- one warm-up round, then one measured round (for `ncu -k regex:gemm2h -s 3 -c 3`)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package(); ops = pkg.ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
R, Cin, Cout = 65536, 1408, 512
x = torch.randn(R, Cin, generator=g).to(dev)
w = (torch.randn(Cout, Cin, generator=g) / Cin ** 0.5).to(dev)
gy = torch.randn(R, Cout, generator=g).to(dev)
for _ in range(2):
    ax, aw, ag = ops._absmax(x), ops._absmax(w), ops._absmax(gy)
    ops._gemm3x(x, False, w, False, R, Cout, Cin, amax_a=ax, amax_b=aw, b_split=ops._wsplit(w, False, aw))
    ops._gemm3x(gy, False, w, True, R, Cin, Cout, amax_a=ag, amax_b=aw, b_split=ops._wsplit(w, True, aw))
    ops._gemm3x(gy, True, x, True, Cout, Cin, R, amax_a=ag, amax_b=ax)
torch.cuda.synchronize()
print("ok")

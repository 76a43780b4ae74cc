"""Does tcgen05.mma kind::tf32 truncate or round fp32 operands whose low 13 mantissa bits are set?"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
ops = pkg.ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
M, N, K = 256, 128, 64
x = torch.randn(M, K, generator=g).to(dev)
w = torch.randn(N, K, generator=g).to(dev)
def trunc(t): return (t.view(torch.int32) & ~0x1fff).view(torch.float32)
def rna(t):
    i = t.view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)
zx, zw = torch.zeros_like(x), torch.zeros_like(w)
wt = trunc(w)
y = ops._gemm3x(x.contiguous(), zx, wt.contiguous(), zw)
yt = trunc(x).double() @ wt.double().t()
yr = rna(x).double() @ wt.double().t()
yf = x.double() @ wt.double().t()
print("err vs trunc(x):", (y.double() - yt).abs().max().item())
print("err vs rna(x):  ", (y.double() - yr).abs().max().item())
print("err vs full x:  ", (y.double() - yf).abs().max().item())
# both operands raw
y2 = ops._gemm3x(x.contiguous(), zx, w.contiguous(), zw)
print("both raw: err vs trunc.trunc:", (y2.double() - trunc(x).double() @ trunc(w).double().t()).abs().max().item(),
      " vs rna.rna:", (y2.double() - rna(x).double() @ rna(w).double().t()).abs().max().item())

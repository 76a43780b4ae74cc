import sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
from oracle import canon, ref_ops as O
pkg = ge.load_package(); dev = torch.device('cuda:0')
def _gen(s): return torch.Generator().manual_seed(s)
def _chunk(B, N, seed):
    g = _gen(seed); side = (N / 4096.0) ** 0.5
    xy = torch.rand(B, N, 2, generator=g) * side + torch.randint(0, 20, (B, 1, 2), generator=g).float()
    z = torch.rand(B, N, 1, generator=g) * 3.0
    return torch.cat((xy, z), dim=2).contiguous()
N, r, K = 24000, 0.1, 32
B, M, D = 2, 1024, 6
xyz = _chunk(B, N, seed=N)
cen = canon.fps(xyz, M, torch.zeros(B, dtype=torch.int32))[1]
o_idx = canon.ball_query(cen, xyz, r, K)
feat = torch.randn(B, N, D, generator=_gen(N + K))
fr = feat.clone().requires_grad_(True)
g_out = O.group(cen, xyz, fr, r, K, True, idx=o_idx.long())
w = torch.randn(g_out.shape, generator=_gen(1))
(g_out * w).sum().backward()
ref64 = torch.zeros(B, N, D, dtype=torch.float64)
for b in range(B):
    ref64[b].index_add_(0, o_idx[b].reshape(-1).long(), w[b].reshape(-1, 3 + D)[:, 3:].double())
for trial in range(3):
    idx = pkg.ops.query_ball_point(r, K, xyz.to(dev), cen.to(dev))
    assert torch.equal(idx.cpu(), o_idx)
    fd = feat.to(dev).requires_grad_(True)
    out = pkg.ops.group_points(xyz.to(dev), fd, cen.to(dev), pkg.ops.NeighborIndex(idx, N), r)
    (out * w.to(dev)).sum().backward()
    gd = fd.grad.cpu()
    diff = (gd - fr.grad).abs()
    bad = ~torch.isclose(gd, fr.grad, rtol=1e-4, atol=1e-5)
    print(f"trial {trial}: max diff {diff.max().item():.3e}, bad {int(bad.sum())}, ours-vs-fp64 {(gd.double()-ref64).abs().max().item():.3e}, torch-vs-fp64 {(fr.grad.double()-ref64).abs().max().item():.3e}")
    for (b, n, d) in bad.nonzero()[:8].tolist():
        cnt = int((o_idx[b] == n).sum())
        print(f"   b={b} n={n} d={d} refs={cnt} ours={gd[b,n,d].item():.7f} torch={fr.grad[b,n,d].item():.7f} fp64={ref64[b,n,d].item():.7f}")

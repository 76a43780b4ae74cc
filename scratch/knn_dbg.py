"""Debug driver for the tensor-core kNN: score bounds, survivors, parity, timing."""
import sys, torch, time
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
from oracle import canon
pkg = ge.load_package(); lib = pkg._lib; lib.load()
dev = torch.device('cuda:0')

def tc_debug(x, k, want_scores=False):
    B, F, N = x.shape
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

for (F, N, offset, scale) in ((64, 512, 0.0, 1.0), (64, 512, 3.0, 0.1), (3, 384, 10.0, 1.0), (20, 640, 0.0, 1.0)):
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, F, N, generator=g) * scale + offset
    idx, scores, stats = tc_debug(x.to(dev), 20, True)
    xd = x.double(); mean = xd.mean(dim=2, keepdim=True)
    v = (x.abs().amax(dim=(1, 2)) + mean.abs().amax(dim=(1, 2)).float()).double()
    S = 2.0 ** (torch.floor(torch.log2(v)) + 1)
    a = (xd - mean) / S.view(-1, 1, 1)
    exact = torch.matmul(a.transpose(1, 2), a) - 0.5 * (a ** 2).sum(1).unsqueeze(1)
    norm = (a ** 2).sum(1).sqrt(); pair = norm.unsqueeze(2) * norm.unsqueeze(1)
    sc = scores.cpu().double()
    gap = exact - sc
    want = canon.knn_expand(x, 20)[0]
    print(f"F={F} N={N} off={offset}: nan={torch.isnan(sc).sum().item()} gap[min,max]=[{gap.min().item():.3e},{gap.max().item():.3e}] "
          f"gap/pair max={(gap / (pair + 1e-12)).max().item():.3e} S={S.tolist()} idx_ok={torch.equal(idx.cpu(), want)} "
          f"mismatch_rows={(idx.cpu() != want).any(-1).sum().item()} stats={stats.tolist()} surv/row={stats[0].item() / (2 * N):.1f}")
    if torch.isnan(sc).any() or gap.abs().max() > 1:
        print("  sample exact", exact[0, 0, :6].tolist()); print("  sample score", sc[0, 0, :6].tolist())
        print("  sample exact r1", exact[0, 1, :6].tolist()); print("  sample score r1", sc[0, 1, :6].tolist())
        print("  col 130..134 exact", exact[0, 0, 130:134].tolist(), "score", sc[0, 0, 130:134].tolist())

# timing at the bench shape
x = torch.randn(16, 4096, 64, device=dev).transpose(1, 2)
for _ in range(3): pkg.ops.knn_graph(x, 20)
torch.cuda.synchronize()
lib.prof_enable(True)
for _ in range(10): pkg.ops.knn_graph(x, 20)
prof = lib.prof_collect(); lib.prof_enable(False)
for k, d in prof.items(): print(f"{k:28s} {1e3 * d['ms'] / d['calls']:9.1f} us/launch  flops/s {d['flops'] / d['ms'] / 1e9:9.1f} T  bytes/s {d['bytes'] / d['ms'] / 1e6:9.1f} GB/s")
idx, _, stats = tc_debug(x, 20)
print("bench-shape survivors/row", stats[0].item() / (16 * 4096), "overflow rows", stats[1].item())

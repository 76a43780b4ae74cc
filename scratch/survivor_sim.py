"""CPU simulation: survivors of the tensor-core kNN filter under different margins (design study)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import ref_ops as O
torch.manual_seed(0)
net = O.DGCNNWithColor(13, k=20, tie="canon")
net.train()
pts, lab, lens = O.s3dis_blocks(2, 4096, 0, 13)
x = pts[:, :, :6].transpose(1, 2)
feats = {}
with torch.no_grad():
    x1 = net.conv1(x[:, :3]); x2 = net.conv2(x1); x3 = net.conv3(x2)
for name, X in (("x1", x1), ("x2", x2), ("x3", x3)):
    xb = X[0].t().double()            # (N,F)
    N = xb.shape[0]
    xc = xb - xb.mean(0)
    nrm = xc.norm(dim=1)
    S = 2 * xc @ xc.t() - (nrm ** 2)[None, :]         # score (row i, col j)
    srt = S.sort(dim=1, descending=True).values
    kth = srt[:, 19]
    # group-max tau: 64 interleaved groups: column j -> group (j % 64)  (approximation of the kernel's grouping)
    G = S.view(N, N // 64, 64).max(dim=1).values       # (N,64)
    tau = G.sort(dim=1, descending=True).values[:, 19]
    for label, cf in (("3xTF32 5e-5", 5e-5), ("1xTF32 4e-3", 4e-3), ("bf16 3.2e-2", 3.2e-2)):
        margin = cf * nrm * nrm.max()
        surv_exact = (S >= (kth - margin)[:, None]).sum(1).float()
        surv_group = (S >= (tau - margin)[:, None]).sum(1).float()
        print(f"{name} {label}: exact-kth thr: mean {surv_exact.mean():.1f} max {surv_exact.max():.0f} | group-max thr: mean {surv_group.mean():.1f} p99 {surv_group.quantile(0.99):.0f} max {surv_group.max():.0f} | |x'| mean {nrm.mean():.2f} max {nrm.max():.2f} d_k^2 mean {(nrm**2 - kth).mean():.3f}")

"""One launch of every hot libpcnbr kernel at the BASELINE shapes (for `ncu --set full`): PointNet++ SA1/FP1 ops at
B=32 and the DGCNN ops at B=16, N=4096, k=20, plus the biggest 1x1-convolution GEMMs."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package(); ops = pkg.ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(reps):
    # ---- PointNet++ SA1 / FP1, B = 32
    B, N, C, K = 32, 4096, 1024, 32
    pts, _, _ = pkg.synthetic.s3dis_blocks(B, N, seed=1)
    xyz = pts[:, :, :3].contiguous().to(dev)
    feat = pts[:, :, 3:].contiguous().to(dev).requires_grad_(True)
    start = torch.zeros(B, dtype=torch.int32, device=dev)
    _, cen = ops.farthest_point_sample(xyz, C, start, return_coords=True)
    nbr = ops.NeighborIndex(ops.query_ball_point(0.1, K, xyz, cen), N)
    grouped = ops.group_points(xyz, feat, cen, nbr, None, pad4=True)
    bn = torch.nn.BatchNorm2d(32).to(dev)
    w = (torch.randn(32, grouped.shape[-1], generator=g) * 0.3).to(dev).requires_grad_(True)
    h = ops.linear_bn_act_rows(grouped, w, None, bn, 0.0)
    bn2 = torch.nn.BatchNorm2d(64).to(dev)
    w2 = (torch.randn(64, 32, generator=g) * 0.2).to(dev).requires_grad_(True)
    pooled = ops.linear_bn_act_maxpool_rows(h, w2, None, bn2, 0.0)          # last SA1 layer fused with the max pool
    pooled.sum().backward()
    i3, d3 = ops.knn_points(xyz, cen, 3)
    n3 = ops.NeighborIndex(i3, C)
    coarse = torch.randn(B, C, 128, generator=g).to(dev).requires_grad_(True)
    up = ops.three_interpolate(coarse, n3, d3)
    up.sum().backward()
    # ---- DGCNN, B = 16
    B, N, k = 16, 4096, 20
    x = torch.randn(B, N, 64, generator=g).to(dev).requires_grad_(True)
    ec = pkg.dgcnn.EdgeConv(64, 64, k).to(dev)
    out = ec(x.transpose(1, 2))
    out.sum().backward()
    rows = torch.randn(B * N, 1408, generator=g).to(dev).requires_grad_(True)
    w6 = (torch.randn(512, 1408, generator=g) * 0.03).to(dev).requires_grad_(True)
    bn6 = torch.nn.BatchNorm1d(512).to(dev)
    y = ops.linear_bn_act_rows(rows, w6, None, bn6, 0.2)
    y.sum().backward()
torch.cuda.synchronize()
print("ok")

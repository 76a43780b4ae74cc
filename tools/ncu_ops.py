"""ONE launch of every hot libpcnbr kernel at the BASELINE shapes, for `ncu --set full` (profiles/README.md):

    ncu --set full --clock-control none --import-source on -k regex:'^(ball_|bn_|csr_|edgeconv|fps_|gemm|grid_|group_|interp|knn_|segsum|select_|sumsq|pool_|maxpool|split_|absmax)' -o gpurun_out/ops python tools/ncu_ops.py [groups]

groups (default all): pnpp = FPS / ball query / 3-NN / group / interpolate / CSR at the PointNet++ SSG shapes (32 x 4096);
dgcnn = feature kNN / CSR / fused EdgeConv at 16 x 4096, F = 64, k = 20; gemm = the conv6 GEMMs (65536 x 1408 -> 512)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
pkg = ge.load_package()
ops = pkg.ops
dev = torch.device("cuda:0")
groups = set(sys.argv[1:]) or {"pnpp", "dgcnn", "gemm"}
g = torch.Generator().manual_seed(0)

if "pnpp" in groups:
    B, N = 32, 4096
    pts, _, _ = pkg.synthetic.s3dis_blocks(B, N, seed=1)
    xyz = pts[:, :, :3].contiguous().to(dev)
    z = torch.zeros(B, dtype=torch.int32, device=dev)
    _, c1 = ops.farthest_point_sample(xyz, 1024, z, return_coords=True)
    _, c2 = ops.farthest_point_sample(c1, 256, z, return_coords=True)
    b1 = ops.NeighborIndex(ops.query_ball_point(0.1, 32, xyz, c1), N)                 # cell grid
    b2 = ops.NeighborIndex(ops.query_ball_point(0.2, 32, c1, c2), 1024)               # scan
    f0 = torch.randn(B, N, 6, generator=g).to(dev)
    f1 = torch.randn(B, 1024, 64, generator=g).to(dev).requires_grad_(True)
    ops.group_points(xyz, f0, c1, b1, None, pad4=True)                                # SA1 gather
    g2 = ops.group_points(c1, f1, c2, b2, None, pad4=True)                            # SA2 gather
    g2.backward(torch.randn(g2.shape, generator=g).to(dev))                           # CSR + segsum<group_bwd>
    i3, d3 = ops.knn_points(xyz, c1, 3)                                               # 3-NN, cell grid
    ff = torch.randn(B, 1024, 128, generator=g).to(dev).requires_grad_(True)
    up = ops.three_interpolate(ff, ops.NeighborIndex(i3, 1024), d3)                   # FP1 interpolation
    up.backward(torch.randn(up.shape, generator=g).to(dev))                           # CSR + segsum<interp_bwd>
    torch.cuda.synchronize()

if "dgcnn" in groups:
    B, N, F, k = 16, 4096, 64, 20
    x = torch.randn(B, N, F, generator=g).to(dev)                                     # point-major features
    nbr = ops.NeighborIndex(ops.knn_graph(x.transpose(1, 2), k), N)
    conv = pkg.dgcnn.EdgeConv(64, 64, k).to(dev)
    xin = x.transpose(1, 2).detach().requires_grad_(True)
    y = conv(xin, _nbr=nbr)
    y.backward(torch.randn(y.shape, generator=g).to(dev))
    torch.cuda.synchronize()

if "gemm" in groups:
    R = 65536
    a = torch.randn(R, 1408, generator=g).to(dev).requires_grad_(True)
    w = (torch.randn(512, 1408, generator=g) * 0.03).to(dev).requires_grad_(True)
    yy = ops.linear_rows(a.view(16, 4096, 1408), w, None)
    yy.backward(torch.randn(yy.shape, generator=g).to(dev))
    n1 = torch.randn(32 * 1024 * 32, 32, generator=g).to(dev).requires_grad_(True)    # SA1 layer 2: 1 M rows x 32 -> 32
    w1 = (torch.randn(32, 32, generator=g) * 0.2).to(dev).requires_grad_(True)
    y1 = ops.linear_rows(n1.view(32, 1024 * 32, 32), w1, None)
    y1.backward(torch.randn(y1.shape, generator=g).to(dev))
    torch.cuda.synchronize()
print("done")

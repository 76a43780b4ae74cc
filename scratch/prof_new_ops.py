"""One launch of the kernels added at the end of round 1 (for `ncu --set full`): multi-radius ball query at the
PointNet++ MSG level-1 shape, the packed block-batch gather and the sliding-window merge."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
ops = pkg.ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, N, C = 32, 4096, 1024
pts, _, _ = pkg.synthetic.s3dis_blocks(B, N, seed=1)
xyz = pts[:, :, :3].contiguous().to(dev)
_, cen = ops.farthest_point_sample(xyz, C, torch.zeros(B, dtype=torch.int32, device=dev), return_coords=True)
for _ in range(2):
    ops.query_ball_point_multi([0.05, 0.1], [16, 32], xyz, cen)
sizes = torch.randint(3000, 12000, (512,), generator=g).tolist()
store = pkg.block_datasets.PackedBlocks([(torch.randn(n, 9, generator=g), torch.zeros(n, 14, dtype=torch.uint8)) for n in sizes], dev)
ids = torch.randint(0, len(store), (256,), generator=g).tolist()
sel = store.draw_device(ids, 4096)
for _ in range(2):
    store.batch(ids, 4096, sel)
scene = torch.randn(1_000_000, 6, generator=g).to(dev)


class Lin(torch.nn.Module):
    num_classes = 13

    def forward(self, x):
        return x.transpose(1, 2).repeat(1, 1, 3)[:, :, :13].contiguous(), None, None


for _ in range(2):
    pkg.dgcnn_utils.predict_single_scene(Lin(), scene, "cuda", 4096, 512)
torch.cuda.synchronize()

"""One FPS launch at the PointNet++ level-1 shape (32 x 4096 -> 1024) for `ncu --set full`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
dev = torch.device("cuda:0")
pts, _, _ = pkg.synthetic.s3dis_blocks(32, 4096, seed=1)
xyz = pts[:, :, :3].contiguous().to(dev)
start = torch.zeros(32, dtype=torch.int32, device=dev)
for _ in range(2):
    pkg.ops.farthest_point_sample(xyz, 1024, start)
torch.cuda.synchronize()

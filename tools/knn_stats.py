import sys, torch
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import __graft_entry__ as ge
from torch.profiler import profile, ProfilerActivity
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
pkg = ge.load_package(); dev = torch.device('cuda:0')
from test_gpu_kernels import _tc_debug
net = pkg.DGCNNWithColor(13, k=20).to(dev)
pts, lab, lens = pkg.synthetic.s3dis_blocks(16, 4096, 0)
x = pts[:, :, :6].transpose(1, 2).to(dev)
acts = {}
for n in ("conv2", "conv3", "conv4"):
    getattr(net, n).register_forward_hook(lambda m, i, o, n=n: acts.__setitem__(n, i[0].detach()))
with torch.no_grad(): net(x)
for n, a in acts.items():
    idx, _, stats = _tc_debug(pkg, a, 20)
    print(n, "shape", tuple(a.shape), "strides", a.stride(), "avg survivors/row %.1f" % (stats[0].item() / (16 * 4096)), "overflow rows", stats[1].item(),
          "mean %.3f std %.3f" % (a.mean().item(), a.std().item()))
xr = torch.randn(16, 64, 4096, device=dev)
idx, _, stats = _tc_debug(pkg, xr, 20)
print("randn avg survivors/row %.1f overflow %d" % (stats[0].item() / (16 * 4096), stats[1].item()))
a = acts["conv2"]
for _ in range(3): pkg.ops.knn_graph(a, 20)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pkg.ops.knn_graph(a, 20); pkg.ops.knn_graph(xr, 20); torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print(f"{e.device_time_total/1e3:8.3f} ms x{e.count} {e.key[:100]}")

"""Does the side-stream geometry of the NEXT batch overlap the captured step?  Kernel timeline (CUPTI via torch.profiler) of two
graph replays of the PointNet++ step with geometry prefetch: per-stream busy time, and what runs while FPS runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import __graft_entry__ as ge
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
pkg = ge.load_package(); dev = torch.device("cuda:0")
model = sys.argv[1] if len(sys.argv) > 1 else "pointnetpp"
B = 32 if model == "pointnetpp" else 16
net = (pkg.PointNetpp(13) if model == "pointnetpp" else pkg.DGCNNWithColor(13, k=20)).to(dev)
bucket = pkg.train.FlatGradBucket(net, steal_grads=True)
opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True, fused=True)
pts, lab, lens = [t.to(dev) for t in pkg.synthetic.s3dis_blocks(B, 4096, 0)]
inp = (lambda p: p) if model == "pointnetpp" else (lambda p: p[:, :, :6].transpose(1, 2))
def loss_of(m, p, l, n, geometry=None):
    out = m(inp(p), geometry=geometry) if geometry is not None else m(inp(p))
    out = out[0] if isinstance(out, tuple) else out
    return pkg.train.masked_onehot_cross_entropy(out, l, n)
geo = lambda m, p, l, n, stream=None: m.prepare_geometry(inp(p), stream=stream)
for _ in range(3):
    bucket.zero(); loss_of(net, pts, lab, lens).backward(); pkg.ops.join_aux(); bucket.all_reduce_mean(); opt.step()
step = pkg.train.GraphedTrainStep(net, opt, bucket, loss_of, (pts, lab, lens), warmup=2, geometry_fn=geo if "--no-prefetch" not in sys.argv else None)
for _ in range(3): step(pts, lab, lens)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(pts, lab, lens); step(pts, lab, lens)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA" and e.device_time_total > 0]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
streams = {}
for e in ev:
    streams.setdefault(getattr(e, "stream", None) if hasattr(e, "stream") else e.device_index, []).append(e)
print("kernels:", len(ev), "span us:", ev[-1].time_range.end - t0)
for s, lst in streams.items():
    print("stream", s, "kernels", len(lst), "busy us", sum(x.time_range.end - x.time_range.start for x in lst))
fps = [e for e in ev if "fps" in e.name]
for f in fps[:4]:
    ov = [e for e in ev if e is not f and e.time_range.start < f.time_range.end and e.time_range.end > f.time_range.start]
    print(f"fps {f.time_range.start - t0:.0f}-{f.time_range.end - t0:.0f} us overlaps {len(ov)} kernels:", sorted({e.name[:28] for e in ov})[:6])

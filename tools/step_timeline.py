"""Kernel timeline (CUPTI via torch.profiler) of ONE replay of the captured train step: span, time with no kernel running
(gaps), the largest gaps with the kernels either side of them, and per-kernel totals inside the graph (warm, in-step).
    python tools/step_timeline.py [dgcnn|pointnetpp] [--top N]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import __graft_entry__ as ge
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
pkg = ge.load_package(); dev = torch.device("cuda:0")
model = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "dgcnn"
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
B = 32 if model == "pointnetpp" else 16
net = (pkg.PointNetpp(13) if model == "pointnetpp" else pkg.DGCNNWithColor(13, k=20)).to(dev)
bucket = pkg.train.FlatGradBucket(net, steal_grads=True)
opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True, fused=True)
pts, lab, lens = [t.to(dev) for t in pkg.synthetic.s3dis_blocks(B, 4096, 0)]
inp = (lambda p: p) if model == "pointnetpp" else (lambda p: p[:, :, :6].transpose(1, 2))
def loss_of(m, p, l, n):
    out = m(inp(p))
    out = out[0] if isinstance(out, tuple) else out
    return pkg.train.masked_onehot_cross_entropy(out, l, n)
for _ in range(3):
    bucket.zero(); loss_of(net, pts, lab, lens).backward(); pkg.ops.join_aux(); bucket.all_reduce_mean(); opt.step()
step = pkg.train.GraphedTrainStep(net, opt, bucket, loss_of, (pts, lab, lens), warmup=2)
for _ in range(5): step(pts, lab, lens)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(pts, lab, lens)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA" and e.device_time_total > 0]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
span = max(e.time_range.end for e in ev) - t0
# union of busy intervals
busy = 0.0; cur_end = t0; gaps = []
last = None
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, last.name if last else "-", e.name))
        busy += t - s; cur_end = t; last = e
    else:
        if t > cur_end:
            busy += t - cur_end; cur_end = t; last = e
print(f"{model}: {len(ev)} kernels, span {span:.0f} us, busy {busy:.0f} us, idle {span - busy:.0f} us in {len(gaps)} gaps")
short = lambda n: n.replace("void ", "").replace("pcnbr::", "")[:60]
hist = collections.Counter()
for g, a, b in gaps: hist[min(int(g), 10)] += 1
print("gap histogram (us: count):", dict(sorted(hist.items())))
for g, a, b in sorted(gaps, reverse=True)[:top]:
    print(f"  {g:6.1f} us  {short(a)}  ->  {short(b)}")
tot = collections.defaultdict(lambda: [0.0, 0])
for e in ev:
    k = short(e.name).split("<")[0].split("(")[0]
    tot[k][0] += e.time_range.end - e.time_range.start; tot[k][1] += 1
print("per kernel (us total, launches):")
for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"  {us:8.1f} {n:4d}  {k}")
if "--seq" in sys.argv:
    print("sequence (start us, dur us, overlaps-previous, name):")
    prev_end = t0
    for e in ev:
        s, t = e.time_range.start, e.time_range.end
        print(f"  {s - t0:8.1f} {t - s:7.1f} {'||' if s < prev_end else '  '} {short(e.name)}")
        prev_end = max(prev_end, t)

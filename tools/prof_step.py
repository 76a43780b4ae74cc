import sys, torch, json
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import __graft_entry__ as ge
from torch.profiler import profile, ProfilerActivity
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
pkg = ge.load_package(); dev = torch.device('cuda:0')
model = sys.argv[1] if len(sys.argv) > 1 else "dgcnn"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
net = (pkg.DGCNNWithColor(13, k=20) if model == "dgcnn" else pkg.PointNetpp(13)).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
pts, lab, lens = pkg.synthetic.s3dis_blocks(B, 4096, 0)
pts, lab, lens = pts.to(dev), lab.to(dev), lens.to(dev)
def step():
    opt.zero_grad(set_to_none=True)
    x = pts[:, :, :6].transpose(1, 2) if model == "dgcnn" else pts
    out = net(x); out = out[0] if isinstance(out, tuple) else out
    loss = pkg.train.masked_onehot_cross_entropy(out, lab, lens); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device ms {tot:.2f}")
for k, ms, c in rows[:40]: print(f"{ms:9.3f} ms  x{c:<4d} {k[:150]}")
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9)

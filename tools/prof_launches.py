"""Per-launch timeline of ONE eager train step (csrc/prof.cu: CUDA events around every libpcnbr kernel, with the launch's
algorithmic bytes / flops), in launch order.  python tools/prof_launches.py dgcnn|pointnetpp|pointnetpp_msg|pointnext [B] [N]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
pkg = ge.load_package(); dev = torch.device('cuda:0')
model = sys.argv[1] if len(sys.argv) > 1 else "dgcnn"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (16 if model == "dgcnn" else 32)
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
net = {"dgcnn": lambda: pkg.DGCNNWithColor(13, k=20), "pointnetpp": lambda: pkg.PointNetpp(13),
       "pointnetpp_msg": lambda: pkg.PointNetppMSG(13), "pointnext": lambda: pkg.PointNeXt(13)}[model]().to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
pts, lab, lens = pkg.synthetic.s3dis_blocks(B, N, 0)
pts, lab, lens = pts.to(dev), lab.to(dev), lens.to(dev)
def step():
    opt.zero_grad(set_to_none=True)
    x = pts[:, :, :6].transpose(1, 2) if model == "dgcnn" else pts
    out = net(x); out = out[0] if isinstance(out, tuple) else out
    loss = pkg.train.masked_onehot_cross_entropy(out, lab, lens); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
lib = pkg._lib.load()
lib.pcnbr_prof_enable(1)
step(); torch.cuda.synchronize()
cap = 1 << 22
buf = ctypes.create_string_buffer(cap)
lib.pcnbr_prof_collect(ctypes.cast(buf, ctypes.c_void_p), cap)
lib.pcnbr_prof_enable(0)
tot = 0.0
print(f"# {model} B={B} N={N}: kernel, us, algorithmic MB, GFLOP, GB/s, TFLOP/s")
for line in buf.value.decode().splitlines():
    name, ms, nb, fl = line.split("\t"); ms = float(ms); nb = float(nb); fl = float(fl); tot += ms
    print(f"{name:42s} {ms*1e3:9.1f} {nb/1e6:10.2f} {fl/1e9:9.2f} {nb/1e9/(ms/1e3):9.0f} {fl/1e12/(ms/1e3):8.1f}")
print(f"# sum of kernel times {tot:.3f} ms")

"""How far is our fp32 PointNet++ from the fp64 twin, relative to the reference's own fp32 evaluation? (per parameter)"""
import sys, os, copy, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import __graft_entry__ as ge
from oracle import ref_ops as O
pkg = ge.load_package()
dev = torch.device("cuda:0")
pts, _, _ = O.s3dis_blocks(2, 4096, seed=0)
torch.manual_seed(3)
ref = O.PointNetpp(13, tie="canon"); ref.drop.p = 0.0
net = pkg.PointNetpp(13); net.drop.p = 0.0
net.load_state_dict(ref.state_dict()); net = net.to(dev)
ref64 = copy.deepcopy(ref).double()
st = torch.tensor([1, 2], dtype=torch.int32)
for name in ("sa1", "sa2", "sa3", "sa4"):
    getattr(net, name).fps_start = st.to(dev)
    getattr(ref, name).fps_start = getattr(ref64, name).fps_start = st
w = torch.randn(2, 4096, 13, generator=torch.Generator().manual_seed(1))
lo = ref(pts); (lo * w).sum().backward()
lo64 = ref64(pts.double()); (lo64 * w.double()).sum().backward()
lg = net(pts.to(dev)); (lg * w.to(dev)).sum().backward()
def rep(name, ours, r32, r64):
    ours, r32, r64 = ours.detach().cpu().double(), r32.detach().double(), r64.detach()
    s = r64.abs().max().item(); er = (r32 - r64).abs().max().item(); eo = (ours - r64).abs().max().item()
    return name, eo / max(er, 1e-300), eo / max(s, 1e-300), er / max(s, 1e-300)
rows = [rep("logits", lg, lo, lo64)]
pr, pr64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
for k, p in net.named_parameters():
    rows.append(rep(k, p.grad, pr[k].grad, pr64[k].grad))
rows.sort(key=lambda r: -r[1])
print("tag", os.environ.get("TAG", ""), "worst ratios (ours_err/ref_err, ours_rel, ref_rel):")
for r in rows[:12]:
    print(f"  {r[0]:40s} {r[1]:8.2f} {r[2]:10.3e} {r[3]:10.3e}")
import statistics
print("  median ratio", statistics.median(r[1] for r in rows), "mean", sum(r[1] for r in rows) / len(rows))

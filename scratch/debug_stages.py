import sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
from oracle import ref_ops as O
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
pkg = ge.load_package(); dev = torch.device('cuda:0')
def err(a, b, name):
    a = a.detach().cpu(); b = b.detach().cpu()
    print(f"{name:28s} shape {tuple(b.shape)} maxerr {(a-b).abs().max().item():.3e} scale {b.abs().max().item():.3e} strides {a.stride()}")
pts, _, _ = O.s3dis_blocks(2, 4096, seed=0)
torch.manual_seed(3)
ref = O.PointNetpp(13, tie="canon"); ref.drop.p = 0.0
net = pkg.PointNetpp(13); net.drop.p = 0.0
net.load_state_dict(ref.state_dict()); net = net.to(dev)
st = torch.tensor([1, 2], dtype=torch.int32)
c0, f0 = pts[:, :, :3], pts[:, :, 3:]
cr, fr = c0, f0
cg, fg = c0.to(dev), f0.to(dev)
feats_r, feats_g, coords_r, coords_g = [f0], [fg], [c0], [cg]
for name in ("sa1", "sa2", "sa3", "sa4"):
    a, b = getattr(net, name), getattr(ref, name)
    a.fps_start, b.fps_start = st.to(dev), st
    # inner stages
    cen_g = pkg.common.sample(cg, a.C, a.fps_start); cen_r = O.sample(cr, b.C, st)
    err(cen_g, cen_r, name + ".sample")
    gg = pkg.common.group(cen_g, cg, fg, a.radius, a.K, a.grouping_norm)
    gr = O.group(cen_r, cr, fr, b.radius, b.K, b.grouping_norm)
    err(gg, gr, name + ".group")
    xg = gg.permute(0, 3, 1, 2); xr = gr.permute(0, 3, 1, 2)
    for li, (conv_g, bn_g, conv_r, bn_r) in enumerate(zip(a.point_net.conv, a.point_net.batch, b.point_net.conv, b.point_net.batch)):
        xg = conv_g(xg); xr = conv_r(xr); err(xg, xr, f"{name}.conv{li}")
        xg = torch.relu(bn_g(xg)); xr = torch.relu(bn_r(xr)); err(xg, xr, f"{name}.bnrelu{li}")
    pg = pkg.common.reduce(xg.permute(0, 2, 3, 1), "max"); pr = O.reduce(xr.permute(0, 2, 3, 1), "max")
    err(pg, pr, name + ".reduce")
    err(pg, xg.permute(0, 2, 3, 1).max(dim=2)[0], name + ".reduce_vs_torch_gpu")
    cg, fg, cr, fr = cen_g, pg, cen_r, pr
    feats_r.append(fr); feats_g.append(fg); coords_r.append(cr); coords_g.append(cg)
# FP
for name, (i1, i2) in (("fp4", (3, 4)), ("fp3", (2, 3)), ("fp2", (1, 2)), ("fp1", (0, 1))):
    a, b = getattr(net, name), getattr(ref, name)
    f1g = feats_g[i1] if name != "fp1" else None; f1r = feats_r[i1] if name != "fp1" else None
    ug = pkg.common.interpolate(feats_g[i2], coords_g[i1], coords_g[i2]); ur = O.interpolate(feats_r[i2], coords_r[i1], coords_r[i2])
    err(ug, ur, name + ".interp")
    og = a(coords_g[i1], coords_g[i2], f1g, feats_g[i2]); orr = b(coords_r[i1], coords_r[i2], f1r, feats_r[i2])
    err(og, orr, name + ".out")
    feats_g[i1], feats_r[i1] = og, orr

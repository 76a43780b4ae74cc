"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not travel to the GPU box):

    CUDA_VISIBLE_DEVICES="" python oracle/make_golden.py

Every fixture stores the inputs and the reference's raw outputs.  Inputs are chosen so that the
reference's torch.topk has no ties among the selected keys (Tier A validity): the
script asserts raw == canonical selection before saving, so the fixtures pin the canonical oracle
and the CUDA path as well.
"""
from __future__ import annotations

import os
import sys

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")   # reference dgcnn.py:39 picks 'cuda' if available
import torch

REF = os.environ.get("PCNBR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from models.utils import common as RC                      # noqa: E402  (reference)
from models.dgcnn import dgcnn as RD                       # noqa: E402  (reference)
from models.PointNetpp.PointNetpp import PointNetpp as RefPointNetpp   # noqa: E402

from oracle import ref_ops as O                            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def filled_cloud(B, N, side, g):
    return torch.rand(B, N, 3, generator=g) * side


def main():
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(20261018)
    fx = {}

    # ---- sample() : S3DIS-shaped block, seeded start draw (common.py:22)
    pts, _, _ = O.s3dis_blocks(2, 512, seed=3)
    xyz = pts[:, :, :3].contiguous()
    torch.manual_seed(77)
    out = RC.sample(xyz, 64)
    torch.manual_seed(77)
    start = torch.randint(0, 512, (2,), dtype=torch.int)
    fx["fps"] = dict(xyz=xyz, C=64, start=start, coords=out)

    # ---- group(): filled cloud -> every r=0.1 ball holds >= K points
    p = filled_cloud(2, 1024, 0.2, g)
    feat = torch.randn(2, 1024, 6, generator=g)
    cen = O.sample(p, 48, torch.tensor([3, 900], dtype=torch.int))
    for norm in (False, True):
        ref = RC.group(cen, p, feat, 0.1, 32, norm)
        assert torch.equal(ref, O.group(cen, p, feat, 0.1, 32, norm, tie="canon")), "ties in group fixture"
        fx[f"group_norm{int(norm)}"] = dict(centroids=cen, coords=p, features=feat, r=0.1, K=32, normalize=norm, out=ref)

    # ---- interpolate()
    coarse = torch.randn(2, 48, 16, generator=g)
    ref = RC.interpolate(coarse, p, cen)
    assert torch.equal(ref, O.interpolate(coarse, p, cen, tie="canon"))
    fx["interpolate"] = dict(points=coarse, coords_1=p, coords_2=cen, out=ref)

    # ---- reduce()
    x = torch.randn(2, 5, 7, 11, generator=g)
    fx["reduce"] = dict(x=x, max=RC.reduce(x, "max"), avg=RC.reduce(x, "avg"))

    # ---- knn(): unit-scale features (distinct distances) in F=3 and F=64
    for Fd, N, k in ((3, 256, 20), (64, 256, 20), (20, 128, 8)):
        xf = torch.randn(2, Fd, N, generator=g)
        ref = RD.knn(xf, k)
        assert torch.equal(ref, O.knn(xf, k, "canon")), "ties in knn fixture"
        fx[f"knn_F{Fd}"] = dict(x=xf, k=k, idx=ref)

    # ---- get_graph_feature() with and without given idx
    xf = torch.randn(2, 8, 128, generator=g)
    ref = RD.get_graph_feature(xf, k=8)
    fx["graph_feature"] = dict(x=xf, k=8, out=ref)

    # ---- EdgeConv module forward (train-mode BN), seeded parameters
    torch.manual_seed(11)
    ec = RD.EdgeConv(8, 16, k=8)
    torch.manual_seed(11)
    ec_o = O.EdgeConv(8, 16, k=8)
    assert all(torch.equal(a, b) for a, b in zip(ec.state_dict().values(), ec_o.state_dict().values()))
    fx["edgeconv"] = dict(x=xf, seed=11, cin=8, cout=16, k=8, out=ec(xf).detach())

    # ---- SetAbstraction + FeaturePropagation module forwards on the filled cloud
    torch.manual_seed(12)
    sa = RC.SetAbstraction(48, 0.1, 9, [16, 16, 32])
    torch.manual_seed(99)                                   # start draw inside sample()
    c1, f1 = sa(p, feat)
    torch.manual_seed(99)
    st = torch.randint(0, 1024, (2,), dtype=torch.int)
    fx["set_abstraction"] = dict(coords=p, features=feat, seed=12, start=st, C=48, r=0.1, cin=9, mlps=[16, 16, 32],
                                 centroids=c1.detach(), out=f1.detach())
    torch.manual_seed(13)
    fp = RC.FeaturePropagation(6 + 32, [32, 16])
    fx["feature_propagation"] = dict(coords_1=p, coords_2=c1.detach(), features_1=feat, features_2=f1.detach(), seed=13,
                                     cin=38, mlps=[32, 16], out=fp(p, c1, feat, f1).detach())

    # ---- PointNet++ SSG logits + gradients on a filled cloud (N=1024), dropout disabled
    N = 1024
    xyzf = filled_cloud(2, N, 0.2, g) + torch.tensor([4.0, 9.0, 0.0])
    rgb = torch.randint(0, 256, (2, N, 3), generator=g).float()
    x9 = torch.cat([xyzf, rgb, xyzf - xyzf.mean(dim=1, keepdim=True)], dim=-1)
    torch.manual_seed(21)
    net = RefPointNetpp(13)
    net.drop.p = 0.0
    torch.manual_seed(21)
    net_o = O.PointNetpp(13, tie="canon")
    net_o.drop.p = 0.0
    sd_r, sd_o = net.state_dict(), net_o.state_dict()
    assert list(sd_r.keys()) == list(sd_o.keys()) and all(torch.equal(sd_r[k], sd_o[k]) for k in sd_r)
    # the reference draws its FPS starts from the global generator: replay the same draws
    torch.manual_seed(314)
    draws = [torch.randint(0, n_src, (2,), dtype=torch.int) for n_src in (N, 1024, 256, 64)]
    torch.manual_seed(314)
    logits = net(x9)
    wgt = torch.randn(2, N, 13, generator=g)
    (logits * wgt).sum().backward()
    for sa_o, st in zip((net_o.sa1, net_o.sa2, net_o.sa3, net_o.sa4), draws):
        sa_o.fps_start = st
    logits_o = net_o(x9)
    assert torch.equal(logits, logits_o), "restated PointNet++ differs from the reference"
    fx["pointnetpp"] = dict(
        x=x9, seed=21, fps_starts=draws, loss_weight=wgt, logits=logits.detach(),
        grads={k: p_.grad.clone() for k, p_ in net.named_parameters()
               if k in ("sa1.point_net.conv.0.weight", "sa4.point_net.conv.2.weight", "fp1.point_net.conv.3.weight", "conv.weight")},
        state_keys=[(k, tuple(v.shape)) for k, v in sd_r.items()],
    )

    # ---- DGCNN / DGCNNWithColor logits + grads (small: N=256, emb 64), dropout disabled
    xd = torch.cat([torch.randn(2, 3, 256, generator=g), torch.rand(2, 3, 256, generator=g)], dim=1)
    for name, Ref, Orc in (("dgcnn", RD.DGCNN, O.DGCNN), ("dgcnn_color", RD.DGCNNWithColor, O.DGCNNWithColor)):
        torch.manual_seed(31)
        m = Ref(num_classes=13, k=20, emb_dims=64, dropout=0.0)
        torch.manual_seed(31)
        mo = Orc(num_classes=13, k=20, emb_dims=64, dropout=0.0)
        sd_r, sd_o = m.state_dict(), mo.state_dict()
        assert list(sd_r.keys()) == list(sd_o.keys()) and all(torch.equal(sd_r[k], sd_o[k]) for k in sd_r)
        lg, emb, _ = m(xd)
        wd = torch.randn(2, 256, 13, generator=g)
        (lg * wd).sum().backward()
        lo = mo(xd)[0]
        assert torch.allclose(lg, lo, rtol=1e-5, atol=1e-6), "restated DGCNN differs from the reference"
        fx[name] = dict(x=xd, seed=31, emb_dims=64, k=20, loss_weight=wd, logits=lg.detach(), emb=emb.detach(),
                        grads={k: p_.grad.clone() for k, p_ in m.named_parameters()
                               if k in ("conv1.conv.0.weight", "conv3.conv.0.weight", "conv8.weight")},
                        state_keys=[(k, tuple(v.shape)) for k, v in sd_r.items()])

    # ---- full-size state_dict key lists (drop-in checkpoint compatibility)
    from models.PointNeXt.PointNeXt import PointNeXt as RefPointNeXt
    fx["state_keys"] = {
        "PointNetpp": [(k, tuple(v.shape)) for k, v in RefPointNetpp(14).state_dict().items()],
        "PointNeXt": [(k, tuple(v.shape)) for k, v in RefPointNeXt(14, "s").state_dict().items()],
        "DGCNN": [(k, tuple(v.shape)) for k, v in RD.DGCNN(13).state_dict().items()],
        "DGCNNWithColor": [(k, tuple(v.shape)) for k, v in RD.DGCNNWithColor(13).state_dict().items()],
    }

    for name, d in fx.items():
        torch.save(d, os.path.join(OUT, f"{name}.pt"))
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(fx)} fixtures, {total / 1024:.0f} KiB -> {OUT}")


if __name__ == "__main__":
    main()

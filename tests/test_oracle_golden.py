"""CPU: the oracle (oracle/canon.c and oracle/ref_ops.py) against the golden vectors produced by the
unmodified reference (oracle/make_golden.py).  When /root/reference is present, also against the
reference imported live."""
import os
import sys

import pytest
import torch

from oracle import canon, ref_ops as O

REF = "/root/reference"
HAVE_REF = os.path.isdir(os.path.join(REF, "models"))


def test_fps_matches_reference_golden(golden):
    g = golden("fps")
    idx, coords = canon.fps(g["xyz"], g["C"], g["start"])
    assert torch.equal(coords, g["coords"])
    assert torch.equal(O.sample(g["xyz"], g["C"], g["start"]), g["coords"])
    assert torch.equal(O.fps_indices(g["xyz"], g["C"], g["start"]), idx)


@pytest.mark.parametrize("name", ["group_norm0", "group_norm1"])
def test_group_matches_reference_golden(golden, name):
    g = golden(name)
    idx = canon.ball_query(g["centroids"], g["coords"], g["r"], g["K"])
    out = canon.group(g["centroids"], g["coords"], g["features"], idx, g["r"], g["normalize"])
    assert torch.equal(out, g["out"])                         # bit-exact, values and order
    assert torch.equal(idx.long(), O.ball_query_indices(g["centroids"], g["coords"], g["r"], g["K"], "canon"))
    assert torch.equal(O.group(g["centroids"], g["coords"], g["features"], g["r"], g["K"], g["normalize"]), g["out"])


def test_interpolate_matches_reference_golden(golden):
    g = golden("interpolate")
    idx, d2 = canon.knn_direct(g["coords_1"], g["coords_2"], 3)
    assert torch.equal(canon.interp(g["points"], idx, d2), g["out"])
    assert torch.equal(O.interpolate(g["points"], g["coords_1"], g["coords_2"]), g["out"])


def test_reduce_matches_reference_golden(golden):
    g = golden("reduce")
    assert torch.equal(O.reduce(g["x"], "max"), g["max"])
    assert torch.equal(O.reduce(g["x"], "avg"), g["avg"])
    with pytest.raises(ValueError):
        O.reduce(g["x"], "sum")


@pytest.mark.parametrize("name", ["knn_F3", "knn_F64", "knn_F20"])
def test_knn_matches_reference_golden(golden, name):
    g = golden(name)
    idx, pd = canon.knn_expand(g["x"], g["k"])
    assert torch.equal(idx.long(), g["idx"])
    assert torch.equal(O.knn(g["x"], g["k"], "canon"), g["idx"])
    # the C restatement reproduces the reference's fp32 distance values bit for bit
    assert torch.equal(pd, torch.gather(O.pairwise_neg_sqdist(g["x"]), 2, g["idx"]))


def test_graph_feature_matches_reference_golden(golden):
    g = golden("graph_feature")
    idx, _ = canon.knn_expand(g["x"], g["k"])
    assert torch.equal(canon.edge_feature(g["x"], idx), g["out"])
    assert torch.equal(O.get_graph_feature(g["x"], g["k"]), g["out"])


def test_sumsq_cascade_order():
    gen = torch.Generator().manual_seed(5)
    for F in (1, 3, 15, 16, 17, 33, 64, 100, 256, 300):
        x = torch.randn(2, F, 70, generator=gen)
        assert torch.equal(canon.sumsq(x), torch.sum(x ** 2, dim=1)), F


def test_canonical_ties_on_lattice():
    """Lattice coordinates j/256: every product/sum is exact, ties are real -> lowest index wins."""
    gen = torch.Generator().manual_seed(9)
    xyz = torch.randint(-20, 21, (2, 300, 3), generator=gen).float() / 256
    q = xyz[:, :40].contiguous()
    idx = canon.ball_query(q, xyz, 0.05, 16)
    assert torch.equal(idx.long(), O.ball_query_indices(q, xyz, 0.05, 16, "canon"))
    i3, d3 = canon.knn_direct(q, xyz, 5)
    d_ref, i_ref = O.three_nn(q, xyz, 5, "canon")
    assert torch.equal(i3.long(), i_ref) and torch.equal(d3, d_ref)
    x = torch.randint(-127, 128, (2, 64, 200), generator=gen).float() / 256
    ik, _ = canon.knn_expand(x, 20)
    assert torch.equal(ik.long(), O.knn(x, 20, "canon"))


def test_underfilled_and_padded_batches():
    """S3DIS-shaped blocks (every ball under-filled), zero padding and duplicates
    (data_processing/block_datasets.py:19-25,122-125)."""
    pts, _, _ = O.s3dis_blocks(2, 700, seed=4)
    xyz = pts[:, :, :3].contiguous()
    xyz[1, 500:] = 0.0                                    # zero padding of a short block
    xyz[0, 600:] = xyz[0, :100]                           # sampling with replacement
    start = torch.tensor([0, 699], dtype=torch.int32)
    idx, cen = canon.fps(xyz, 64, start)
    assert torch.equal(idx, O.fps_indices(xyz, 64, start))
    b = canon.ball_query(cen, xyz, 0.1, 32)
    assert torch.equal(b.long(), O.ball_query_indices(cen, xyz, 0.1, 32, "canon"))


def test_models_match_reference_golden(golden):
    g = golden("pointnetpp")
    torch.manual_seed(g["seed"])
    net = O.PointNetpp(13)
    net.drop.p = 0.0
    for sa, st in zip((net.sa1, net.sa2, net.sa3, net.sa4), g["fps_starts"]):
        sa.fps_start = st
    logits = net(g["x"])
    assert torch.allclose(logits, g["logits"], rtol=1e-4, atol=1e-5)
    (logits * g["loss_weight"]).sum().backward()
    params = dict(net.named_parameters())
    for k, v in g["grads"].items():
        assert torch.allclose(params[k].grad, v, rtol=1e-3, atol=1e-4), k
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [(k, tuple(s)) for k, s in g["state_keys"]]

    for name, cls in (("dgcnn", O.DGCNN), ("dgcnn_color", O.DGCNNWithColor)):
        g = golden(name)
        torch.manual_seed(g["seed"])
        m = cls(num_classes=13, k=g["k"], emb_dims=g["emb_dims"], dropout=0.0)
        logits, emb, _ = m(g["x"])
        assert torch.allclose(logits, g["logits"], rtol=1e-4, atol=1e-5)
        assert torch.allclose(emb, g["emb"], rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
def test_oracle_against_live_reference():
    sys.path.insert(0, REF)
    try:
        from models.utils import common as RC
        from models.dgcnn import dgcnn as RD
    finally:
        sys.path.remove(REF)
    gen = torch.Generator().manual_seed(123)
    xyz = torch.rand(2, 600, 3, generator=gen) * torch.tensor([1.0, 1.0, 3.0]) + torch.tensor([13.0, 2.0, 0.0])
    torch.manual_seed(8)
    ref = RC.sample(xyz, 50)
    torch.manual_seed(8)
    start = torch.randint(0, 600, (2,), dtype=torch.int)
    assert torch.equal(canon.fps(xyz, 50, start)[1], ref)
    x = torch.randn(2, 64, 300, generator=gen)
    assert torch.equal(canon.knn_expand(x, 20)[0].long(), RD.knn(x, 20))
    feats = torch.randn(2, 50, 24, generator=gen)
    i3, d3 = canon.knn_direct(xyz, ref, 3)
    assert torch.equal(canon.interp(feats, i3, d3), RC.interpolate(feats, xyz, ref))


def test_metrics_match_reference_golden(golden):
    """Training/metrics.py: the oracle's restatement against the unmodified reference's outputs
    (oracle/make_golden_metrics.py), incl. an argmax tie, a padded and an empty cloud."""
    g = golden("metrics")
    cm = O.metrics_confusion_matrix(g["pred"], g["labels"], g["mask"])
    assert torch.equal(cm, g["confusion"])
    assert O.metrics_update_accuracy(g["pred"], g["labels"], g["mask"]) == (g["correct"], g["total"])
    inter, union = O.metrics_update_iou(g["pred"], g["labels"], g["mask"])
    assert torch.equal(inter, g["inter"]) and torch.equal(union, g["union"])
    miou, ious = O.metrics_iou(g["pred"], g["labels"], g["mask"])
    assert torch.equal(ious, g["ious"]) and abs(miou - g["miou"]) < 1e-7


def test_msg_composition_matches_reference_functions(golden):
    """BASELINE configs[2]: the oracle's multi-scale set abstraction against the composition of the unmodified
    reference's sample / group / MiniPointNet / reduce (oracle/make_golden_msg.py)."""
    g = golden("msg")
    for r, K, want in zip(g["radii"], g["Ks"], g["tables"]):
        got = O.ball_query_indices(g["centroids"], g["coords"], r, K, tie="canon")
        assert torch.equal(got.to(torch.int32), want)
        assert torch.equal(canon.ball_query(g["centroids"], g["coords"], r, K), want)       # the C restatement
    torch.manual_seed(g["seed"])
    m = O.SetAbstractionMSG(g["C"], g["radii"], g["cin"], g["mlps"], g["Ks"])
    m.fps_start = g["start"]
    cen, out = m(g["coords"], g["features"])
    assert torch.equal(cen, g["centroids"]) and torch.equal(out.detach(), g["out"])


# --------------------------------------------------------------------------- round-2 fixtures (oracle/make_golden_large.py)

def _chunk(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    side = (N / 4096.0) ** 0.5
    xy = torch.rand(B, N, 2, generator=g) * side + torch.randint(0, 20, (B, 1, 2), generator=g).float()
    z = torch.rand(B, N, 1, generator=g) * 3.0
    return torch.cat((xy, z), dim=2).contiguous()


def test_knn_baseline_shape_matches_reference_golden(golden):
    """F=64, N=4096, k=20 (BASELINE configs[1]): the C oracle equals the unmodified reference's knn() indices."""
    g = golden("knn_F64_N4096")
    x = torch.randn(1, 64, 4096, generator=torch.Generator().manual_seed(int(g["seed"])))
    assert torch.equal(canon.knn_expand(x, int(g["k"]))[0].to(torch.int16), g["idx"])


def test_large_fixtures_match_reference_golden(golden):
    g = golden("fps_24k")
    xyz = _chunk(1, g["N"], g["seed"])
    assert torch.equal(canon.fps(xyz, g["C"], g["start"])[1], g["coords"])
    g = golden("group_8k")
    gen = torch.Generator().manual_seed(g["seed"])
    p = torch.rand(1, 8192, 3, generator=gen) * 0.5
    feat = torch.randn(1, 8192, 6, generator=gen)
    idx = canon.ball_query(g["centroids"], p, g["r"], g["K"])
    assert torch.equal(canon.group(g["centroids"], p, feat, idx, g["r"], True), g["out"])
    h = golden("interp_8k")
    i3, d3 = canon.knn_direct(p, g["centroids"], 3)
    out = canon.interp(h["points"], i3, d3)
    assert torch.equal(out[:, :512], h["out_first"]) and torch.equal(out.double().sum(dim=1), h["out_sum"])


def test_invresmlp_and_pointnext_oracle_match_reference_golden(golden):
    g = golden("invresmlp")
    gen = torch.Generator().manual_seed(g["data_seed"])
    pc = torch.rand(2, g["N"], 3, generator=gen) * 0.15 + torch.tensor([3.0, 8.0, 0.0])
    f = torch.randn(2, g["N"], 64, generator=gen)
    torch.manual_seed(g["seed"])
    blk = O.InvResMLP(g["radius"], g["cin"], g["width"], g["K"], tie="canon")
    assert torch.equal(blk(pc, pc, f)[1], g["out"])


# --------------------------------------------------------------------------- length-aware fixtures (oracle/make_golden_lengths.py)

def test_per_cloud_oracle_matches_reference_golden_on_unpadded_clouds(golden):
    """The checker of the length-aware forms is the C oracle applied to each cloud's real rows; the fixtures are the
    unmodified reference run on each cloud alone.  They must agree before the CUDA path is compared with either."""
    g = golden("lengths")
    L, xyz = g["lengths"].tolist(), g["xyz"]
    for b, n in enumerate(L):
        p = xyz[b:b + 1, :n].contiguous()
        cen = canon.fps(p, g["sample"]["C"], g["sample"]["start"][b:b + 1])[1]
        assert torch.equal(cen[0], g["sample"]["coords"][b])
        gr = g["group"]
        idx = canon.ball_query(cen, p, gr["r"], gr["K"])
        assert torch.equal(canon.group(cen, p, gr["features"][b:b + 1, :n].contiguous(), idx, gr["r"], True)[0], gr["out"][b])
        i3, d3 = canon.knn_direct(p, cen, 3)
        assert torch.equal(canon.interp(g["interpolate"]["points"][b:b + 1], i3, d3)[0], g["interpolate"]["out"][b, :n])
        for F in (3, 64):
            k = g[f"knn_F{F}"]
            assert torch.equal(canon.knn_expand(k["x"][b:b + 1, :, :n].contiguous(), k["k"])[0][0], k["idx"][b, :n])
